"""Pin a timing dependence of the pair kernel to a hand-off: runs one case of tests/chaos_case.py under random warp delays
restricted to subsets of the trace sites (B200_MLA_CHAOS_SITES bit mask: bit 4 * role + event for the pipeline's sites, bits 20-25 the
request-boundary detail sites) and prints, per mask, how many runs differ from the undisturbed output and where.
Usage: python tools/chaos_sites.py [case index] [repeats] [mask ...]"""
import os, sys, subprocess, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
case = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
masks = [int(m, 0) for m in sys.argv[3:]] or [0xfffff, 0x3ffffff]
code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import chaos_case as C
C.CASES = [C.CASES[%d]]
C.main(sys.argv[1], int(sys.argv[2]))
''' % (ROOT, os.path.join(ROOT, "tests"), case)
env = dict(os.environ); env.pop("B200_MLA_TRACE", None)
subprocess.run([sys.executable, "-c", code, "/tmp/plain.npz", "1"], env=env, check=True, capture_output=True)
p = np.load("/tmp/plain.npz")
for m in masks:
    env["B200_MLA_TRACE"] = os.environ.get("CHAOS_TRACE", "5"); env["B200_MLA_CHAOS_SITES"] = str(m)
    r = subprocess.run([sys.executable, "-c", code, "/tmp/chaos.npz", str(reps)], env=env, capture_output=True, text=True)
    if r.returncode != 0:
        print("mask 0x%07x: run failed: %s" % (m, r.stderr[-400:])); continue
    c = np.load("/tmp/chaos.npz")
    bad = 0
    for key in c.files:
        base = key.rsplit("_", 1)[0] + "_0"
        if np.array_equal(c[key], p[base]):
            continue
        bad += 1
        a, b = c[key], p[base]
        if key.startswith("o_"):
            a = (a.view(np.uint16).astype(np.uint32) << 16).view(np.float32); b = (b.view(np.uint16).astype(np.uint32) << 16).view(np.float32)
            d = np.abs(a.astype(np.float64) - b.astype(np.float64)).reshape(a.shape[0], -1, a.shape[-1])
            for req in np.unique(np.argwhere(d > 0)[:, 0]):
                rows = np.where((d[req] > 0).any(axis=1))[0]; cols = np.where((d[req] > 0).any(axis=0))[0]
                # contiguous column runs
                runs = []; s0 = cols[0]; prev = cols[0]
                for cc in cols[1:]:
                    if cc != prev + 1: runs.append((int(s0), int(prev))); s0 = cc
                    prev = cc
                runs.append((int(s0), int(prev)))
                print("  mask 0x%07x %s request %d: %d rows [%d..%d], %d cols in runs %s, max %.4f" % (m, key, req, len(rows), rows[0], rows[-1], len(cols), runs[:8], d[req].max()))
        else:
            d = np.abs(a.astype(np.float64) - b.astype(np.float64)); d[np.isnan(d)] = np.inf
            idx = np.argwhere((d > 0) | (np.isnan(a) != np.isnan(b)))
            print("  mask 0x%07x %s differs: shape %s, %d elements, max %.3e, first %s: %s vs %s" % (m, key, a.shape, len(idx), d.max(), idx[:4].tolist(), [float(a[tuple(i)]) for i in idx[:4]], [float(b[tuple(i)]) for i in idx[:4]]))
    print("mask 0x%07x: %d of %d outputs differ" % (m, bad, len(c.files)))
print("done")
