"""Compose profiles/r02_scaling_summary.json from the per-N bench lines and H2D probes already under profiles/:
python scripts/scaling_summary.py > profiles/r02_scaling_summary.json"""
import json, os, sys
P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
rows = []
for n in (1, 2, 4, 8):
    b = json.load(open(os.path.join(P, "r02_bench_final.json" if n == 1 else f"r02_bench_n{n}.json")))
    h = json.load(open(os.path.join(P, f"r02_h2d_n{n}.json")))
    gbs_min = h["min_per_gpu_gbs"]["fp32_batch_192MiB"]
    batch = b["config"].get("batch_per_gpu", 16384)
    bytes_per_step = b["e2e"]["h2d_bytes_per_step"]
    ceiling = n * batch / (bytes_per_step / (gbs_min * 1e9))  # timing is max over ranks: the slowest GPU's copy sets it
    u8 = b if n == 1 else json.load(open(os.path.join(P, f"r02_bench_n{n}_rerun.json")))  # final uint8 chunk plan
    rows.append({"n_gpus": n, "value": b["value"], "e2e": b["e2e"]["value"], "e2e_u8": u8["e2e_u8"]["value"],
                 "rerun_other_allocation": None if n == 1 else {"value": u8["value"], "e2e": u8["e2e"]["value"],
                                                                "source": f"profiles/r02_bench_n{n}_rerun.json"},
                 "parity_bit_exact": b["parity"]["bit_exact"],
                 "h2d_aggregate_gbs_fp32_batch": h["aggregate_gbs"]["fp32_batch_192MiB"], "h2d_min_per_gpu_gbs": gbs_min,
                 "h2d_min_per_gpu_gbs_uint8_batch": h["min_per_gpu_gbs"]["uint8_batch_48MiB"],
                 "fp32_e2e_ceiling_max_over_ranks": ceiling,
                 "source": "profiles/" + ("r02_bench_final.json" if n == 1 else f"r02_bench_n{n}.json")})
for r in rows:
    for k in ("value", "e2e", "e2e_u8"):
        r["efficiency_" + k] = r[k] / (r["n_gpus"] * rows[0][k])
    r["e2e_over_its_ceiling"] = r["e2e"] / r["fp32_e2e_ceiling_max_over_ranks"]
print(json.dumps({"what": "weak scaling of bench.py at N = 1/2/4/8 GPUs of one box (separate gpurun allocations, all on the final "
                          "code), with the concurrent pinned H2D bandwidth measured in the same allocation (N=1: an earlier "
                          "allocation); ceiling = images/s the slowest GPU's H2D rate allows (timing is max over ranks)",
                  "rows": rows}, indent=1))
