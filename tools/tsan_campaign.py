"""Seeded campaign of the ThreadSanitizer race check (tests/jit_tsan.py) over the kernels of random plans: circuits,
planner switches, dtypes, grids, tile blocks, load forms, chunk-sized launches.  CPU only.
    python tools/tsan_campaign.py [first_seed] [n_seeds] [workers]"""
import os
import sys
from concurrent.futures import ProcessPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def one(seed: int):
    import numpy as np
    from quantum_simulations_b200 import workloads as W
    from quantum_simulations_b200.kernel.cuda_dense import compile_circuit
    from tests.jit_tsan import race_check

    rng = np.random.default_rng(seed)
    n = int(rng.integers(14, 17))
    dtype = "complex64" if rng.random() < 0.3 else "complex128"
    os.environ["QSV_JIT_PAIR"] = str(int(rng.integers(0, 3)))
    kind = int(rng.integers(0, 3))
    cd = (W.random_1q_cz(n, int(rng.integers(6, 22)), seed) if kind == 0 else
          W.random_mixed(n, int(rng.integers(60, 260)), seed) if kind == 1 else W.qft(n))
    kw = {}
    if rng.random() < 0.3:
        kw["low_store_round"] = False
    if rng.random() < 0.3:
        kw["park_reorder"] = True
    if rng.random() < 0.3:
        kw["warp_local_rounds"] = True
        os.environ["QSV_JIT_WARP_SYNC"] = "1"
    else:
        os.environ.pop("QSV_JIT_WARP_SYNC", None)
    if rng.random() < 0.4:
        kw.update(explore_seed=int(rng.integers(0, 1000)), explore_k=int(rng.choice([3, 5])))
    prog = compile_circuit(cd, dtype=dtype, zero_init=bool(rng.random() < 0.5), **kw)
    checked = 0
    for k, step in enumerate(prog.passes[:2]):
        n_tiles = (1 << n) >> 11
        grid, blk = int(rng.integers(1, 4)), int(rng.integers(0, 3))
        tr = (0, 1) if (step.desc.zero_input and rng.random() < 0.5) else \
             (int(rng.integers(0, 3)), int(rng.integers(n_tiles - 3, n_tiles + 1))) if rng.random() < 0.4 else None
        reports, text = race_check(step, n, dtype=dtype, grid=grid, tile_block=blk, tile_range=tr)
        if reports:
            return seed, f"FAIL n={n} {dtype} pair={os.environ['QSV_JIT_PAIR']} pass={k} grid={grid} blk={blk} range={tr} kw={kw}\n{text[-1500:]}"
        checked += 1
    return seed, f"ok {checked} kernels n={n} {dtype} pair={os.environ['QSV_JIT_PAIR']}"


if __name__ == "__main__":
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    workers = int(sys.argv[3]) if len(sys.argv) > 3 else max(1, (os.cpu_count() or 2) // 2)
    bad = total = 0
    with ProcessPoolExecutor(max_workers=workers) as ex:
        for seed, msg in ex.map(one, range(first, first + count)):
            print(seed, msg.splitlines()[0], flush=True)
            if msg.startswith("FAIL"):
                print(msg)
            bad += msg.startswith("FAIL")
            total += int(msg.split()[1]) if msg.startswith("ok") else 0
    print(f"tsan campaign: {count} plans, {total} kernels race-checked, {bad} with reports")
    sys.exit(1 if bad else 0)
