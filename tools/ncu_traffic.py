"""DRAM traffic of the dominant kernel from an `ncu --set full` capture -> profiles/gemm_tc_traffic.json (read by bench.py).

  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel -s 38 -c 8 \
      -o gpurun_out/prof_gemm_rX -f python tools/prof_step.py fp16 64
  python tools/ncu_traffic.py gpurun_out/prof_gemm_rX.ncu-rep

With -s 38 the eight captured launches are two consecutive Swin stage-3 blocks at batch 64 (qkv, proj, fc1, fc2: M = 36864,
C = 768), whose algorithmic bytes are stated next to the measured ones."""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
def col(name): return hdr.index(name)
def to_bytes(v, u):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    return float(v) * f
M, C = 64 * 576, 768
shapes = [("s3.qkv", M, 3 * C, C, 2, False), ("s3.proj", M, C, C, 4, True), ("s3.fc1", M, 4 * C, C, 2, False), ("s3.fc2", M, C, 4 * C, 4, True)]
launches = []
for i, r in enumerate(rows[2:]):
    rd = to_bytes(r[col("dram__bytes_read.sum")], units[col("dram__bytes_read.sum")])
    wr = to_bytes(r[col("dram__bytes_write.sum")], units[col("dram__bytes_write.sum")])
    name, m, n, k, ob, res = shapes[i % 4]
    algo = (m * k + n * k) * 2 + m * n * ob + (m * n * 4 if res else 0) + (m * n * 2 if res else 0)   # + the raw 16-bit rows the folded LayerNorm consumes
    launches.append(dict(launch=name, M=m, N=n, K=k, dram_read=rd, dram_write=wr, dram_total=rd + wr, algorithmic_bytes=algo,
                         traffic_over_algorithmic=(rd + wr) / algo, duration_us=float(r[col("gpu__time_duration.sum")]),
                         tensor_pipe_pct=float(r[col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")])
                         if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in hdr else None))
out = dict(kernel="gemm_tc_kernel", source=os.path.basename(rep), launches_captured=len(launches),
           dram_bytes_per_launch=sum(l["dram_total"] for l in launches) / max(1, len(launches)),
           algorithmic_bytes_per_launch=sum(l["algorithmic_bytes"] for l in launches) / max(1, len(launches)),
           note="two consecutive Swin stage-3 blocks at batch 64 (fp16): qkv, proj(+fp32 residual), fc1(GELU), fc2(+fp32 residual)",
           launches=launches)
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "profiles", "gemm_tc_traffic.json"), "w"), indent=1)
for l in launches:
    print(f"{l['launch']:8s} {l['duration_us']:7.1f} us  dram {l['dram_total']/1e6:7.1f} MB  algorithmic {l['algorithmic_bytes']/1e6:7.1f} MB  x{l['traffic_over_algorithmic']:.2f}  tensor {l['tensor_pipe_pct']}")
