"""Reads the bench lines of tools/gpu_r2_ab_loads.sh, prints the table, and writes the environment of the fastest
variant whose final state has the baseline's fingerprint (stdout: shell exports; table on stderr)."""
import json
import sys
from pathlib import Path

d = Path(sys.argv[1])
rows = []
for f in sorted(d.glob("bench_blk*_pair*.json")):
    try:
        r = json.loads(f.read_text().strip().splitlines()[-1])
    except Exception as e:
        print(f"{f.name}: no line ({e})", file=sys.stderr)
        continue
    blk, pair = f.stem.split("_")[1][3:], f.stem.split("_")[2][4:]
    c = r["config"]
    rows.append({"blk": int(blk), "pair": int(pair), "ms": r["ms_per_step"], "frac": r["roofline"]["frac"],
                 "per_pass_ms": c["per_pass_ms"], "fp": c.get("state_fingerprint")})
    print(f"blk={blk} pair={pair}  {r['ms_per_step']:.2f} ms  frac {r['roofline']['frac']:.3f}  {c['per_pass_ms']}  fp {c.get('state_fingerprint')}",
          file=sys.stderr)
base = next((r for r in rows if r["blk"] == 0 and r["pair"] == 0), None)
best = base
for r in rows:
    if base is not None and r["fp"] != base["fp"]:
        print(f"blk={r['blk']} pair={r['pair']}: FINGERPRINT DIFFERS from the baseline", file=sys.stderr)
        continue
    if best is None or r["ms"] < best["ms"] * 0.99:
        best = r
(d / "ab_table.json").write_text(json.dumps({"rows": rows, "picked": best}, indent=1))
if best is not None:
    print(f"export QSV_JIT_TILE_BLOCK={best['blk']}; export QSV_JIT_PAIR={best['pair']}")
    print(f"picked blk={best['blk']} pair={best['pair']}", file=sys.stderr)
