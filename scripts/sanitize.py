"""Small ladder for compute-sanitizer (memcheck / racecheck / initcheck are slow: keep shapes tiny)."""
import itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import ascii_chat_b200 as acb
import oracle_bind as ob
assert acb.lib().acb200_init(0) == 0
bad = 0
img = ob.gen("noise", 64, 48, 1); img[:, 20:30] = 0
for level, mode in itertools.product((0, 1, 2, 3), (0, 1, 2)):
    got = acb.ascii_convert_with_capabilities(img, 16, 8, acb.make_caps(level, mode), False, False, "blocks")
    bad += got != ob.port_convert(img, 16, 8, level, mode, "blocks")
img = ob.gen("bars", 320, 96, 2)
for level, mode in ((0, 0), (2, 0), (3, 0), (3, 2), (1, 2)):
    cfg = acb.make_cfg(320, 96, 40, 24 if mode == 2 else 12, level, mode, "standard", scale=acb.SCALE_BOX, pad_left=2, pad_top=1)
    got = acb.render_batch_host(cfg, [img, img[::-1].copy(), img])
    exp = ob.port_convert(img, 40, 12, level, mode, "standard", scale=ob.SCALE_BOX)
    exp = ob._take(ob.port().orc_pad_height(ob._take(ob.port().orc_pad_width(exp, 2)), 1))
    bad += got[0] != exp or got[2] != exp
srcs = [ob.port_convert(ob.gen("noise", 96, 64, i), 20, 6, 3, 0) for i in range(4)]
bad += acb.ascii_create_grid(srcs, 60, 20) != ob.port_create_grid(srcs, 60, 20)
got, _, _ = acb.composite([ob.gen("bars", 80, 60, i) for i in range(3)], 60, 20)
bad += not np.array_equal(got, ob.port_composite([ob.gen("bars", 80, 60, i) for i in range(3)], 60, 20)[0])
print("sanitize ladder mismatches:", bad)
