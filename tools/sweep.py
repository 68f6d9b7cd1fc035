#!/usr/bin/env python
"""Run bench.py over every BASELINE.json workload (configs 2-5) and print / save a table.

  python tools/sweep.py [--out gpurun_out/sweep.jsonl] [--steps 400] [--only substr]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--only", default="")
    ap.add_argument("--cpu-seconds", type=float, default=3.0)
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    rows = []
    with open(a.out, "w") as f:
        for name in bench.WORKLOADS:
            if a.only and a.only not in name:
                continue
            solver = bench.WORKLOADS[name]["prob"] in ("sokoban", "ddave", "mdungeon")
            steps = max(50, a.steps // 4) if solver else a.steps
            cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", name, "--steps", str(steps),
                   "--warmup", "128", "--chunk", "128", "--cpu-seconds", str(a.cpu_seconds)]
            p = subprocess.run(cmd, capture_output=True, text=True)
            line = p.stdout.strip().split("\n")[-1] if p.stdout.strip() else ""
            try:
                d = json.loads(line)
            except Exception:
                print(name, "FAILED", p.stderr[-400:])
                continue
            d["workload"] = name
            f.write(json.dumps(d) + "\n")
            f.flush()
            rows.append(d)
            print("%-28s value %.3e  e2e %.3e  e2e_rollout %.3e  cpu(port,%d cores) %.3e  roofline %.4f" % (
                name, d["value"], d["e2e"]["value"], d.get("e2e_rollout", {}).get("value", float("nan")),
                d["cpu_baseline"]["cores"], d["cpu_baseline"]["value"], d["roofline"]["frac"]), flush=True)


if __name__ == "__main__":
    main()
