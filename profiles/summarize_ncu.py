"""Turn `ncu --set full` reports (scratch, gpurun_out/*.ncu-rep) into the small JSON summaries committed here.
   python profiles/summarize_ncu.py gpurun_out/conv_r01c.ncu-rep gpurun_out/wgrad_r01c.ncu-rep > profiles/<name>.json
Needs the `ncu` CLI (reads the report; no GPU)."""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "time_ns",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct_active",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum": "l2_to_sm_tma_bytes",
    "l1tex__m_xbar2l1tex_read_bytes_mem_dshared.sum": "dsmem_read_bytes",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_bytes_total",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__registers_per_thread": "regs",
    "launch__shared_mem_per_block_dynamic": "smem_dyn_bytes",
    "launch__cluster_size": "cluster_size",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
}


def summarize(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        rec = {"kernel": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]}
        for k, name in KEEP.items():
            if k in ix and r[ix[k]] != "":
                v = float(r[ix[k]].replace(",", ""))
                u = units[ix[k]]
                if u in ("Kbyte", "KB"):
                    v *= 1e3
                elif u in ("Mbyte", "MB"):
                    v *= 1e6
                elif u in ("Gbyte", "GB"):
                    v *= 1e9
                elif u in ("us", "usecond"):
                    v *= 1e3
                elif u in ("ms", "msecond"):
                    v *= 1e6
                rec[name] = v
        out.append(rec)
    return out


if __name__ == "__main__":
    print(json.dumps({p.split("/")[-1]: summarize(p) for p in sys.argv[1:]}, indent=1))
