"""gpurun_out/r2_k_gemm_{fwd,dgrad,wgrad_acc}<suffix>.ncu-rep -> profiles/<out>.json: duration, DRAM bytes, tensor-pipe
activity of one launch each, with the algorithmic bytes beside them (what bench.py reports as roofline.traffic).
    python tools/ncu_gemm_json.py <suffix> <tokens> <out.json>"""
import csv
import json
import subprocess
import sys

suffix, T, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
d, f = 4096, 11008
METRICS = "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
ALG = {  # A + B + D (+ C for the accumulating wgrad), bytes
    "fwd_gateup": 2 * T * d + 2 * 2 * f * d + 2 * T * 2 * f,
    "dgrad_gateup": 2 * T * 2 * f + 2 * 2 * f * d + 2 * T * d,
    "wgrad_gateup_acc": 2 * T * 2 * f + 2 * T * d + 2 * 4 * 2 * f * d,
}
FILES = {"fwd_gateup": "gemm_fwd", "dgrad_gateup": "gemm_dgrad", "wgrad_gateup_acc": "gemm_wgrad_acc"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ns": 1e-3, "ms": 1e3, "%": 1.0}
res = {}
for key, stem in FILES.items():
    rep = f"gpurun_out/r2_k_{stem}{suffix}.ncu-rep"
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", METRICS], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    get = lambda m: float(vals[hdr.index(m)].replace(",", "")) * UNIT[units[hdr.index(m)]]
    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    res[key] = dict(kernel=vals[hdr.index("Kernel Name")][:60], duration_us=get("gpu__time_duration.sum"), dram_read_bytes=rd,
                    dram_write_bytes=wr, algorithmic_bytes=ALG[key],
                    tensor_pipe_active_pct=get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                    traffic_over_algorithmic=round((rd + wr) / ALG[key], 3))
json.dump(dict(source=f"ncu --set full --clock-control none, one launch each at the Llama-2-7B gate|up shapes with M = {T} tokens "
                      f"(tools/one_kernel.py TOKENS={T})", tokens=T, kernels=res), open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
