"""One rank of a Morton-domain run at benchmark-like sizes (developer check, not a test module).

  python scripts/dd_scale_rank.py RANK WORLD COMM DEVICE N STEPS TOKEN OUT.json [DECOMP]

COMM = host (virtual ranks may share one GPU) | nccl.  Prints / writes per-rank domain statistics (own, halo, LET nodes),
stage times of the last step, and the global state fingerprint sums so that a WORLD-rank run can be held against the
1-rank run of the same device-generated disc.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, comm, device, n, steps, token, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), sys.argv[7], sys.argv[8]
    decomp = int(sys.argv[9]) if len(sys.argv) > 9 else 1
    from summersph_b200 import default_params, MODE_VARIABLE_H
    from summersph_b200.engine import Engine
    p = default_params(MODE_VARIABLE_H)
    p.n_ranks = world
    p.decomposition = decomp if world > 1 else 0
    with Engine(p, device=device) as e:
        if world > 1:
            if comm == "nccl":
                e.comm_init(rank, world, bytes.fromhex(token))
            else:
                e.comm_init_host(rank, world, token)
        e.ics_disc(n, seed=20251018)
        dt, t = 0.01, 0.0
        stages = []
        for _ in range(steps):
            e.timer_start()
            dt, t = e.step(dt, t)
            ms = e.timer_stop()
            st = e.stage_times(); st["step"] = ms
            stages.append({k: round(v, 3) for k, v in st.items()})
        h, sums = e.state_hash()
        e.set_exact_counters(True); e.evaluate(); c = e.counters(); e.set_exact_counters(False)
        res = {"rank": rank, "world": world, "n": n, "dt": dt, "t": t, "sizes": list(e.sizes()), "hash": f"{h:016x}", "sums": sums,
               "domain": e.domain_stats(), "stages": stages, "counters": c}
    with open(out, "w") as f:
        json.dump(res, f)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
