"""Summarise an ncu report (one or more kernels) as JSON: python scripts/ncu_summary.py rep.ncu-rep > profiles/x.json
Per launch: kernel name, grid/block, duration, DRAM bytes read+written, tensor-pipe / issue utilisation, top stalls."""
import csv, json, subprocess, sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "sm__cycles_elapsed.max": "sm_cycles",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active": "adu_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid", "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


STAGES = ("quant_conv1", "conv2_pool", "conv3", "conv4_pool", "conv5", "conv6_pool", "fc1", "fc2_dequant")


def main(path, batch=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    ix = {k: i for i, k in enumerate(head)}
    res = []
    for r in rows[2:]:
        d = {"kernel": r[ix["Kernel Name"]]}
        for k, name in KEYS.items():
            if k in ix and r[ix[k]] not in ("", "n/a"):
                v = float(r[ix[k]].replace(",", ""))
                u = units[ix[k]].split("/")[0]
                if name.endswith("_bytes") or name == "duration_us" or name == "dynamic_smem":
                    v *= UNIT_SCALE.get(u, 1.0)
                d[name] = v
        stalls = {k.split("issue_stalled_")[1].split("_per_issue")[0]: float(r[i]) for k, i in ix.items()
                  if "issue_stalled_" in k and k.endswith("per_issue_active.ratio") and r[i] not in ("", "n/a")}
        d["top_stalls"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:4])
        if "dram_read_bytes" in d:
            d["dram_bytes"] = d["dram_read_bytes"] + d.get("dram_write_bytes", 0.0)
        res.append(d)
    if batch is not None and len(res) == len(STAGES):  # the 8 kernels of one forward, in launch order
        res = {"batch": int(batch), "source": path.split("/")[-1], "stages": dict(zip(STAGES, res))}
    json.dump(res, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
