import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1]); print(f, "value %.4g e2e %.4g ms/step %.2f launch_ms %.3f share %.3f clk %s"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["kernel_share_of_step"], d["clocks"]["sm_mhz"]))
    except Exception as e: print(f, "ERR", e)
