import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = sys.argv[1] if len(sys.argv) > 1 else "4096"
for f4, f2 in [(None, None), (0, 1.0), (0.03, 0.97), (0.2, 0.8)]:
    env = dict(os.environ)
    if f4 is not None: env.update(UALM_F4=str(f4), UALM_F2=str(f2))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_profile_dev.py"), B], env=env, capture_output=True, text=True).stdout
    line = [l for l in out.splitlines() if "kernel" in l]
    print("B", B, "F4", f4, "F2", f2, "->", line[0] if line else out[-300:])
