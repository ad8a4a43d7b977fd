"""Which outputs differ between an undisturbed run and runs with random warp delays (tests/chaos_case.py under
B200_MLA_TRACE=5)?  Prints the tensors, how many elements and which (request, head) rows differ.  Usage: python tools/chaos_diff.py"""
import os, sys, subprocess, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
helper = os.path.join(ROOT, "tests", "chaos_case.py")
env = dict(os.environ); env.pop("B200_MLA_TRACE", None)  # (B200_MLA_DEFER_RESCALE=1 in the environment tests that instantiation)
subprocess.run([sys.executable, helper, "/tmp/plain.npz", "1"], env=env, check=True, capture_output=True)
env["B200_MLA_TRACE"] = "5"
subprocess.run([sys.executable, helper, "/tmp/chaos.npz", sys.argv[1] if len(sys.argv) > 1 else "12"], env=env, check=True, capture_output=True)
p, c = np.load("/tmp/plain.npz"), np.load("/tmp/chaos.npz")
for key in c.files:
    base = key.rsplit("_", 1)[0] + "_0"
    if not np.array_equal(c[key], p[base]):
        a, b = c[key], p[base]
        if key.startswith("o_"):
            a = a.view(np.uint16).astype(np.uint32) << 16; b = b.view(np.uint16).astype(np.uint32) << 16
            a = a.view(np.float32); b = b.view(np.float32)
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        idx = np.argwhere(d > 0)
        print(key, "shape", a.shape, "n diff", len(idx), "max", d.max(), "first idx", idx[:3].tolist(), "rows touched (b, s/h..):", sorted(set(map(tuple, idx[:, :3].tolist())))[:12])
print("done")
