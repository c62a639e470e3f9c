#!/bin/bash
# r02k visit (1 GPU): warp-per-line digital rain
TAG=r02k
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -30 $O/${TAG}_pytest.txt | cut -c1-600
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/${TAG}_smoke.txt
echo "== fuzz 90 s"; timeout 400 python scripts/fuzz_parity.py 90 2>&1 | tail -8 | tee $O/${TAG}_fuzz_parity.txt
echo "== rain timing"; timeout 300 python - <<'PY' 2>&1 | tee $O/${TAG}_rain_timing.txt
import sys, time
sys.path[:0] = [".", "tests"]
import ascii_chat_b200 as acb, oracle_bind as ob
assert acb.lib().acb200_init(0) == 0
for (W, H, c, r, level, mode) in ((3840, 2160, 320, 96, 3, 2), (1920, 1080, 160, 48, 3, 0), (640, 480, 80, 24, 0, 0)):
    s = ob.port_convert(ob.gen("noise", W, H, 1), c, r, level, mode)
    rain, ref = acb.DigitalRain(c, r, 3), (ob.RefRain if ob.ref() is not None else ob.PortRain)(c, r, 3)
    ok = all(rain.apply(s, 0.016) == ref.apply(s, 0.016) for _ in range(3))
    t0 = time.perf_counter()
    for _ in range(50):
        out = rain.apply(s, 0.016)
    ours = (time.perf_counter() - t0) / 50 * 1e3
    t0 = time.perf_counter()
    for _ in range(5):
        ref.apply(s, 0.016)
    theirs = (time.perf_counter() - t0) / 5 * 1e3
    print("%dx%d level %d mode %d: in %d B out %d B  ours %.3f ms  reference %.3f ms  identical %s" % (c, r, level, mode, len(s), len(out), ours, theirs, ok))
PY
echo "== sanitizer"; bash scripts/gpu_sanitize.sh 2>&1 | tee $O/${TAG}_compute_sanitizer.txt
ls -la $O | tail -4
