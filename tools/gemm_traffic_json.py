"""profiles/r01_gemm_traffic.json from an ncu --set full capture of tools/gemm_probe.py (B=55, ITERS=1, launch order
qkv, proj, fc1, fc2): per launch the FLOPs, duration under ncu, DRAM traffic and the algorithmic operand/result bytes.
usage: python tools/gemm_traffic_json.py gpurun_out/x.ncu-rep profiles/r01_gemm_traffic.json [B]"""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 55
M = B * 257
shapes = (("qkv N=3072 K=1024 (fp16 out)", 3072, 1024, 2), ("proj N=1024 K=1024 (residual)", 1024, 1024, 8),
          ("fc1 N=4096 K=1024 (QuickGELU)", 4096, 1024, 2), ("fc2 N=1024 K=4096 (residual)", 1024, 4096, 8))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3}


def val(r, name):
    i = hdr.index(name)
    return float(r[i].replace(",", "")) * scale.get(units[i], 1)


res = []
for r, (name, N, K, out_bytes) in zip(rows[2:6], shapes):
    t = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    res.append({"kernel": name, "M": M, "N": N, "K": K, "flops": 2 * M * N * K, "time_us": t,
                "tflops_under_ncu": 2 * M * N * K / t / 1e6, "dram_read_bytes": rd, "dram_write_bytes": wr,
                "traffic_bytes": rd + wr, "algorithmic_bytes": M * K * 2 + N * K * 2 + M * N * out_bytes // (2 if out_bytes == 8 else 1),
                "tensor_pipe_active_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                "registers": val(r, "launch__registers_per_thread")})
json.dump({"source": f"ncu --set full --clock-control none, tools/gemm_probe.py B={B} ({rep})", "launches": res}, open(out, "w"), indent=1)
for d in res:
    print(d["kernel"], f'{d["time_us"]:.1f} us', f'{d["tflops_under_ncu"]:.0f} TF', f'traffic {d["traffic_bytes"] / 1e6:.0f} MB', f'alg {d["algorithmic_bytes"] / 1e6:.0f} MB')
