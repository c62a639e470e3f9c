"""GPU-side diagnosis aid: run a ladder of small cases through the C ABI and, for every mismatch against the
oracle, print where the byte strings first differ.  Output is meant to be read offline (gpurun_out/)."""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import ascii_chat_b200 as acb  # noqa: E402
import oracle_bind as ob  # noqa: E402


def show(tag, got, exp):
    if got == exp:
        return 0
    if got is None:
        print("MISMATCH %s: got None err=%s" % (tag, acb.last_error()))
        return 1
    n = min(len(got), len(exp))
    i = next((k for k in range(n) if got[k] != exp[k]), n)
    print("MISMATCH %s: len got=%d exp=%d first diff @%d\n   got=%r\n   exp=%r" % (
        tag, len(got), len(exp), i, got[max(0, i - 24): i + 40], exp[max(0, i - 24): i + 40]))
    return 1


def main():
    assert acb.lib().acb200_init(0) == 0, acb.last_error()
    bad = tot = 0
    for pat, (W, H, c, r) in itertools.product(("noise", "bars", "solid"), ((8, 2, 8, 2), (64, 48, 16, 8), (320, 240, 80, 24))):
        img = ob.gen(pat, W, H, 3)
        for level, mode in itertools.product((0, 1, 2, 3), (0, 1, 2)):
            for pal in ("standard", "blocks"):
                got = acb.ascii_convert_with_capabilities(img, c, r, acb.make_caps(level, mode), False, False, pal)
                exp = ob.port_convert(img, c, r, level, mode, pal)
                tot += 1
                b = show("nn %s %dx%d->%dx%d L%d M%d %s" % (pat, W, H, c, r, level, mode, pal), got, exp)
                bad += b
                if bad > 12:
                    print("too many mismatches, stopping")
                    return
    for (W, H, c, r) in ((64, 64, 16, 8), (640, 480, 80, 24), (333, 127, 47, 13)):
        img = ob.gen("noise", W, H, 1)
        for level, mode in ((0, 0), (3, 0), (3, 2), (2, 2)):
            cfg = acb.make_cfg(W, H, c, r * 2 if mode == 2 else r, level, mode, "standard", scale=acb.SCALE_BOX)
            try:
                got = acb.render_batch_host(cfg, [img])[0]
            except RuntimeError as e:
                got = None
                print("box error", e)
            exp = ob.port_convert(img, c, r, level, mode, "standard", scale=ob.SCALE_BOX)
            tot += 1
            bad += show("box %dx%d->%dx%d L%d M%d" % (W, H, c, r, level, mode), got, exp)
    print("diag: %d cases, %d mismatches" % (tot, bad))


if __name__ == "__main__":
    main()
