"""dev: compare the scale paths (parts vs one thread per element) at several n."""
import os, sys, subprocess
sys.path.insert(0, ".")
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from ripp_b200 import _lib, synth
    ctx = _lib.Context(0)
    for n in (64, 100, 1000, 4096, 16384):
        for grp in (1, 2):
            for gen in (True, False):
                sc = ctx.to_device(synth.scalars_mont("dbg-%d" % grp, n))
                w = 24 if grp == 1 else 48
                pts = None
                if not gen:
                    pts = (synth.g1_points_dev if grp == 1 else synth.g2_points_dev)(ctx, "dbg-p", n)
                out = ctx.alloc(n * w * 4)
                (ctx.g1_scale_dev if grp == 1 else ctx.g2_scale_dev)(pts, sc, n, out)
                ctx.sync()
                np.save("/tmp/dbg_%s_%d_%d_%d.npy" % (sys.argv[2], n, grp, gen), out.download((n, w)))
else:
    for tag, env in (("parts", {}), ("thread", {"RIPP_B200_SCALE_PARTS_MAX": "0"})):
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, __file__, "child", tag], env=e, check=True)
    for n in (64, 100, 1000, 4096, 16384):
        for grp in (1, 2):
            for gen in (1, 0):
                a = np.load("/tmp/dbg_parts_%d_%d_%d.npy" % (n, grp, gen)); b = np.load("/tmp/dbg_thread_%d_%d_%d.npy" % (n, grp, gen))
                bad = np.nonzero((a != b).any(axis=1))[0]
                print("n=%d G%d gen=%d: %d mismatching elements%s" % (n, grp, gen, len(bad), (" first %s" % bad[:8]) if len(bad) else ""))
