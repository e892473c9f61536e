#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of the library build that is on disk.

    ncu --set full --clock-control none -k regex:'k_aev_forward_cta|k_aev_backward|k_mlp_fused|k_gemm_tc' -c 16 \
        -o gpurun_out/traffic_cap -f python bench.py --steps 1 --warmup 3 --cpu-steps 0 [--config NAME]
    ncu -i gpurun_out/traffic_cap.ncu-rep --page raw --csv > gpurun_out/traffic_cap.csv
    python tools/make_traffic.py gpurun_out/traffic_cap.csv water10k

Adds {library sha256[:16] and build id (torchani_b200.build.build_id): {config: {"mlp": bytes, "aev_forward": bytes, "aev_backward": bytes, "launches": {...}}}} --
dram__bytes_read.sum + dram__bytes_write.sum PER STEP'S LAUNCHES of each kernel family (one step = the last complete
set in the capture; ncu flushes caches between kernels, so these are cold-cache figures).  bench.py reports them as
roofline.traffic only when the hash of the library it runs matches."""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, config = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    i_name = hdr.index("Kernel Name")
    i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    i_t = hdr.index("gpu__time_duration.sum")
    units = rows[1]

    def to_bytes(v, u):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        return float(v.replace(",", "")) * mult

    fam = {"k_mlp_fused": "mlp", "k_gemm_tc": "mlp", "k_aev_forward": "aev_forward", "k_aev_backward": "aev_backward"}
    per = {}
    for r in rows[2:]:
        if len(r) <= max(i_r, i_w):
            continue
        key = next((v for k, v in fam.items() if k in r[i_name]), None)
        if key is None:
            continue
        per.setdefault(key, []).append((to_bytes(r[i_r], units[i_r]) + to_bytes(r[i_w], units[i_w]), r[i_name].split("(")[0]))
    out = {"launches": {}}
    for key, lst in per.items():
        # launches of one step: 1 (AEV kernels, data-flow MLP) or 6 (chained MLP); take the last complete set
        n = 6 if (key == "mlp" and "k_gemm_tc" in lst[-1][1]) else 1
        out[key] = sum(b for b, _ in lst[-n:])
        out["launches"][key] = {"kernel": lst[-1][1], "per_step": n, "captured": len(lst)}
    from torchani_b200 import _lib
    h = hashlib.sha256(open(_lib.LIB_PATH, "rb").read()).hexdigest()[:16]
    dst = os.path.join(ROOT, "profiles", "traffic.json")
    data = json.load(open(dst)) if os.path.exists(dst) else {}
    data.setdefault(h, {})[config] = out
    from torchani_b200 import build as _build
    bid = _build.build_id()          # the same capture under the reproducible identities of the build
    for key in (bid, _build.build_id(kernels_only=True)):
        if key:
            data.setdefault(key, {})[config] = out
    json.dump(data, open(dst, "w"), indent=1)
    print(h, bid, config, out)


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
