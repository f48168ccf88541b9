"""GPU: the reference-side adapter (integration/trieste_b200_adapter.py — the file a trieste maintainer would add) executed
for real.  TensorFlow / trieste cannot be installed here, so the two imports it makes are satisfied by minimal stand-ins: a
`tensorflow` module with `constant` and tensors that have `.numpy()`, and `trieste.models.interfaces` whose Protocol classes
are GENERATED from the committed `ast` fixture of the reference's own file (tests/golden/reference_protocols.json), abstract
methods included — so instantiating the adapter proves it implements every abstract method of the protocols it claims."""
import abc
import importlib.util
import json
import os
import sys
import types

import numpy as np
import pytest

from oracle import gp_oracle as o

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Tensor(np.ndarray):
    def numpy(self):
        return np.asarray(self)


def _install_stubs(monkeypatch):
    tf = types.ModuleType("tensorflow")
    tf.constant = lambda x, dtype=None: np.asarray(x).view(_Tensor)
    tf.Tensor = _Tensor
    monkeypatch.setitem(sys.modules, "tensorflow", tf)
    fixture = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_protocols.json")))["models/interfaces.py"]
    mod = types.ModuleType("trieste.models.interfaces")
    built = {}

    def build(name):
        if name in built:
            return built[name]
        spec = fixture[name]
        bases = tuple(build(b) for b in spec["bases"] if b in fixture) or (abc.ABC,)
        ns = {}
        for mname, m in spec["methods"].items():
            if mname.startswith("__"):
                continue
            src = f"def {mname}(self, {', '.join(m['args'])}):\n    raise NotImplementedError\n"
            loc = {}
            exec(src, {}, loc)
            ns[mname] = abc.abstractmethod(loc[mname])
        built[name] = abc.ABCMeta(name, bases, ns)
        return built[name]

    for name in fixture:
        setattr(mod, name, build(name))
    for pkg in ("trieste", "trieste.models"):
        monkeypatch.setitem(sys.modules, pkg, types.ModuleType(pkg))
    monkeypatch.setitem(sys.modules, "trieste.models.interfaces", mod)
    return tf


def test_reference_side_adapter_runs_against_the_native_engine(monkeypatch):
    tf = _install_stubs(monkeypatch)
    spec = importlib.util.spec_from_file_location("trieste_b200_adapter", os.path.join(ROOT, "integration", "trieste_b200_adapter.py"))
    adapter = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(adapter)

    om = o.synthetic_model(o.hartmann_6, 300, 6)

    class Matern52:  # what the adapter reads from a gpflow GPR
        variance, lengthscales = tf.constant(om.variance), tf.constant(om.lengthscales)

    gpr = types.SimpleNamespace(data=(tf.constant(om.X), tf.constant(om.y)), kernel=Matern52(),
                                mean_function=types.SimpleNamespace(c=tf.constant(om.mean_const)),
                                likelihood=types.SimpleNamespace(variance=tf.constant(om.noise)))
    model = adapter.B200GaussianProcessRegression(gpr)  # abstract methods of all claimed protocols are implemented
    Xq = tf.constant(np.random.default_rng(1).uniform(size=(500, 6)))
    mean, var = model.predict(Xq)
    omean, ovar = o.predict(om, np.asarray(Xq))
    assert hasattr(mean, "numpy")
    np.testing.assert_allclose(mean.numpy(), omean, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(var.numpy(), ovar, rtol=0, atol=1e-9 * om.variance)
    jm, jc = model.predict_joint(tf.constant(np.asarray(Xq)[:40].reshape(8, 5, 6)))
    _, ojc = o.predict_joint(om, np.asarray(Xq)[:40].reshape(8, 5, 6))
    np.testing.assert_allclose(jc.numpy(), ojc, rtol=0, atol=1e-9 * om.variance)
    assert model.sample(Xq[:7][None], 3).shape == (1, 3, 7, 1)
    model.update(types.SimpleNamespace(query_points=tf.constant(om.X[:280]), observations=tf.constant(om.y[:280])))
    model.optimize(None)
    assert len(model.get_internal_data()) == 280
    # the fused acquisition path on the adapter's native model, driven through the reference-shaped optimiser signature
    import trieste_b200 as tb
    from trieste_b200.acquisition import ExpectedImprovement
    from trieste_b200.acquisition.optimizer import generate_random_search_optimizer

    fn = ExpectedImprovement().prepare_acquisition_function(model.native, tb.Dataset(om.X[:280], om.y[:280]))
    pt = generate_random_search_optimizer(20000)(tb.Box([0.0] * 6, [1.0] * 6), fn)
    assert pt.shape == (1, 6)
