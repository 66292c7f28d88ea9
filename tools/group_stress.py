"""Stress of the one-proof-over-a-context-group path: many proofs on groups of 2 / 4 / 8 ranks (one GPU listed several
times, or real GPUs when present), coset-resident FRI on and off, host-descriptor and resident entry points; every proof is
compared with the single-GPU bytes.  usage: group_stress.py [iterations]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
m = importlib.import_module("streaming-zero-knowledge-proofs_b200")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
single = m.Context()
ngpu = torch.cuda.device_count()
bad = 0
total = 0
for world in (8, 4, 2):
    devs = list(range(world)) if ngpu >= world else [0] * world
    cases = []
    for T, b, tau in [(1 << 17, 512, 2), (1 << 19, 512, 3), (1 << 18, 256, 8)]:
        ct = m.simulate(T, b, tau, seed=31 + T % 7)
        root = m.manifest_root(ct)
        cases.append((ct, root, single.prove_v1(ct, root)))
    for it in range(iters):
        g = m.Context(devices=devs)  # fresh group every iteration: cold buffers, tables, events
        for mode in (1, 0, 1):
            g.set_option("fri_coset", mode)
            for ct, root, want in cases:
                got = g.prove_v1(ct, root)
                rt = g.upload_trace(ct)
                got2 = g.prove_v1_resident(rt, root)
                rt.free()
                total += 2
                if got != want or got2 != want:
                    bad += 1
                    print("MISMATCH world", world, "iter", it, "mode", mode, "T", ct.n_rows, "tau", ct.tau, got == want, got2 == want, flush=True)
        g.close()
    print("world", world, "devices", devs, "done", flush=True)
print("proofs", total, "mismatches", bad)
sys.exit(1 if bad else 0)
