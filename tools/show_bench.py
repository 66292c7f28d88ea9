#!/usr/bin/env python
"""Pretty-print the interesting parts of a bench.py JSON line."""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print(f"value {d['value']:.4g} {d['unit']}  ms/step {d['ms_per_step']:.2f}  e2e {d['e2e']['value']:.4g} ({d['e2e']['ms_per_step']:.2f} ms)  launches {d['gpu_launches']}")
print("phases", {k: round(v, 2) for k, v in d["phases_ms"].items()})
r = d["roofline"]
ia = r["int_alu"]
print(f"roofline {r['kernel']}: {r['ms_per_launch']:.2f} ms, {r['achieved']:.1f} GB/s, frac {r['frac']:.4f}, share {r['share_of_step']:.2f}, "
      f"{ia.get('compressions_per_s', ia.get('plain_compressions_per_s', 0)):.3g} compr/s = {ia.get('frac_of_alu_pipe_bound', ia.get('plain_frac_of_alu_pipe_bound', 0)):.3f} of ALU-pipe bound")
if "column_commit" in r:
    c = r["column_commit"]
    print(f"column commit: tables {c['ms_per_launch']:.2f} ms ({c['achieved']:.0f} GB/s, frac {c['frac']:.3f}, {c['tabled_columns']} cols tabled, "
          f"{c['chunks_redone_by_generic_kernel']} chunks redone), dedup only {c['dedup_only_ms_per_launch']:.2f} ms, plain {c['plain_ms_per_launch']:.2f} ms")
if d.get("micro"):
    print("micro", {k: (round(v["ms"], 3), round(v["GBps"], 1), round(v["frac_of_hbm_peak"], 4)) for k, v in d["micro"].items()})
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"].get("gpu_proof_identical"), "clocks", d["clocks"])
if d.get("sharded_single_proof"):
    sh = d["sharded_single_proof"]
    print("group prove", sh["n_gpus"], "GPUs: e2e", round(sh["e2e_ms_per_proof"], 2), "ms, resident", round(sh["resident_ms_per_proof"], 2), "ms, identical",
          sh["identical_to_single_gpu_proof"], {k: round(v, 2) for k, v in sh["resident_phases_ms_gpu0"].items()})
if d.get("sharded_single_proof_nccl_callbacks"):
    sh = d["sharded_single_proof_nccl_callbacks"]
    print("nccl-callback prove", round(sh["ms_per_proof"], 2), "ms e2e,", round(sh["resident"]["ms_per_proof"], 2), "ms resident, identical", sh["identical_on_all_ranks"])
if d.get("worst_case_step"):
    print("worst case step", round(d["worst_case_step"]["ms_per_step"], 2), "ms, identical", d["worst_case_step"]["identical_proof"])
if d.get("lde_commit_fri"):
    lc = d["lde_commit_fri"]
    print("lde_commit_fri", lc["n_gpus"], "GPUs:", round(lc["ms_per_step"], 1), "ms", round(lc["GBps_per_gpu"], 1), "GB/s/GPU alg", f"{lc['compressions_per_s_per_gpu']:.3g} compr/s/GPU",
          "oracle roots", lc["first_columns_match_oracle_digests"], {k: round(v, 1) for k, v in lc["phases_ms_gpu0"].items()})
if d.get("jsonl_stream"):
    j = d["jsonl_stream"]
    print("jsonl_stream", f"{j['rows_per_s']:.4g} rows/s", round(j["ms_per_step"], 1), "ms", round(j["file_MBps"]), "MB/s", j["parser_threads"], "threads",
          "parse", round(j["jsonl_parse_ms"], 1), "ms hidden", round(j["stream_copy_hidden_frac"], 3), "prove", round(j["prove_ms_after_last_line"], 2),
          "ms  python reader", f"{j['python_reader_rows_per_s']:.3g} rows/s", "identical", j["proof_identical_to_one_shot"])
