import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = sys.argv[1] if len(sys.argv) > 1 else "1024"
combos = [(None, None), (0.05, 0.95), (0.1, 0.9), (0.15, 0.85), (0.2, 0.8), (0.0, 1.0), (0.3, 0.7)]
for f4, f2 in combos:
    env = dict(os.environ)
    if f4 is not None: env.update(UALM_F4=str(f4), UALM_F2=str(f2))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_profile_dev.py"), B], env=env, capture_output=True, text=True).stdout
    line = [l for l in out.splitlines() if "kernel" in l]
    print("F4", f4, "F2", f2, "->", line[0] if line else out[-300:])
