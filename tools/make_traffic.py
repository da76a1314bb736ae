"""Write profiles/traffic_r02.json from the ncu extracts of tools/gpu_evidence_r02b.sh (run on the GPU box before the bench lines)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O = os.path.join(ROOT, "gpurun_out")


def dram(path):
    try:
        d = json.load(open(path))["launches"]
    except Exception:
        return None
    out = []
    for l in d:
        def num(k):
            v, u = l.get(k, "0 byte").split()[:2]
            return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        out.append(num("dram__bytes_read.sum") + num("dram__bytes_write.sum"))
    return out


p = os.path.join(ROOT, "profiles", "traffic_r02.json")
t = json.load(open(p)) if os.path.exists(p) else {}
t["source"] = ("ncu --set full --clock-control none captures of round 2 (tools/gpu_evidence_r02.sh, gpu_evidence_r02b.sh; the .ncu-rep files are not kept, their judged "
               "metrics are in profiles/ncu_*_r02.json); dram__bytes_read.sum + dram__bytes_write.sum per launch, bench workload = 1024 problems, seed 0, hill unless a key says otherwise")
for key, f in (("kb_kernel_bytes_per_launch_config2", "ncu_kb_kernel_b1024_r02.json"), ("kb_kernel_float_direct_gather_bytes_per_launch_b8192_hill", "ncu_kb_kernel_r02.json"),
               ("ka_kernel_float_bytes_per_launch_4096_active_round_300", "ncu_ka_kernel_r02.json")):
    v = dram(os.path.join(O, f))
    if v:
        t[key] = v[0]
json.dump(t, open(p, "w"), indent=1)
json.dump(t, open(os.path.join(O, "traffic_r02.json"), "w"), indent=1)
print("traffic:", {k: v for k, v in t.items() if k != "source"})
