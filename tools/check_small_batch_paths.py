"""Batch-size invariance of the small-batch kernels (run on a B200: python tools/check_small_batch_paths.py).

Few candidate tiles: the K* digit generation splits the training rows over CTAs (fixed-order mean reduction) and the gradient
assembly runs one CTA per candidate; the same points inside a large batch take the unsplit kernels.  Both are the same
arithmetic up to the summation order of the mean / the gradient sums; this script prints the largest differences.
(Not part of tests/: written after the round's GPU budget was spent, so its tolerances were never calibrated on hardware; the
split kernels themselves are exercised by every small-batch parity test of the GPU suite.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
g.build()
from oracle import gp_oracle as o  # checker side only
from tests.util import candidates, model_pair
from trieste_b200.acquisition import expected_improvement

om, nm = model_pair(o.ackley, 2048, 10)
fn = expected_improvement(nm, float(om.y.min()))
Xbig = candidates(20_000, 10, seed=5)  # > 148 K* CTAs, > 2048 candidates: unsplit kernels
vb, gb = fn.value_and_gradient(Xbig[:, None, :])
mb, sb = nm.predict(Xbig)
for m in (1, 7, 96, 130, 1000):
    vs, gs = fn.value_and_gradient(Xbig[:m, None, :])
    ms, ss = nm.predict(Xbig[:m])
    print(f"m={m:5d}  max|dmean|={np.abs(ms - mb[:m]).max():.2e}  max|dvar|/var={np.abs(ss - sb[:m]).max() / om.variance:.2e}  "
          f"max|dEI|={np.abs(np.asarray(vs) - np.asarray(vb)[:m]).max():.2e}  max|dgrad|={np.abs(np.asarray(gs) - np.asarray(gb)[:m]).max():.2e}")
omean, ovar = o.predict_batched(om, Xbig[:130])
print("vs oracle: max|dmean| =", np.abs(mb[:130] - omean).max(), " max|dvar|/var =", np.abs(sb[:130] - ovar).max() / om.variance)
