"""profiles/r2_ncu_traffic.json (what bench.py prints as roofline.traffic / whole_frame_dram_bytes) from the `ncu --set full`
captures of one C2 frame: dram__bytes_read.sum + dram__bytes_write.sum per kernel launch.
    python tools/ncu_traffic.py gpurun_out/r2s2_ncu_hybrid_final.ncu-rep gpurun_out/r2s2_ncu_fused.ncu-rep > profiles/r2_ncu_traffic.json"""
import csv, io, json, subprocess, sys


def launches(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for r in data:
        def val(name):
            i = hdr.index(name)
            v = float(r[i].replace(",", ""))
            return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}[units[i]]
        out.append((r[hdr.index("Kernel Name")].split("(")[0].replace("void ", ""), int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))))
    return out


hyb, fus = launches(sys.argv[1]), launches(sys.argv[2])
trace_like = ("k_trace", "k_phong", "k_frame")
res = {"c2": {
    "source": {"hybrid": sys.argv[1], "fused": sys.argv[2], "note": "ncu --set full --clock-control none, one launch per kernel of one frame; caches are flushed before every replay, so these are cold-cache upper bounds"},
    "hybrid_per_kernel_dram_bytes": {k: b for k, b in hyb},
    "hybrid_trace_kernels_dram_bytes": sum(b for k, b in hyb if k.startswith(trace_like)),
    "hybrid_frame_dram_bytes": sum(b for k, b in hyb),
    "k_frame_dram_bytes": sum(b for k, b in fus if k.startswith("k_frame")),
    "frame_dram_bytes": sum(b for k, b in fus),
}}
print(json.dumps(res, indent=1))
