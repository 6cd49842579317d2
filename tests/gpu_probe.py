"""Run every kernel parity case in its own subprocess (timeout each) and write gpurun_out/probe.json.
A faulting or hanging kernel is reported without taking the other cases down.  Usage: python tests/gpu_probe.py [pattern]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def all_cases():
    import graph_cases as G
    import kernel_cases as K
    d = dict(K.CASES)
    d.update(G.CASES)
    d.update(K.PENDING_CASES)          # written but not yet run on a B200: only selected by `--pending` or by name
    d.update(G.PENDING_CASES)
    d.update(G.PROBE_CASES)           # report-only measurements: only selected by `--probe` or by name
    return d


def run_one(name):
    import torch
    torch.cuda.init()
    r = all_cases()[name]()
    print("RESULT " + json.dumps(r))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_one(sys.argv[2])
        sys.exit(0)
    pat = sys.argv[1] if len(sys.argv) > 1 else ""
    cases = all_cases()
    import graph_cases as G
    import kernel_cases as K
    pending = dict(K.PENDING_CASES)
    pending.update(G.PENDING_CASES)
    pending.update(G.PROBE_CASES)
    if pat == "--pending":
        pat = ",".join(k for k in pending if k not in G.PROBE_CASES)
    elif pat == "--probe":
        pat = ",".join(G.PROBE_CASES)
    elif not pat:
        cases = {k: v for k, v in cases.items() if k not in pending}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    report = {}
    for name in cases:
        if pat and not any(q in name for q in pat.split(",")):
            continue
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "--one", name], capture_output=True, text=True, timeout=int(os.environ.get("UB_CASE_TIMEOUT", "300")))
            res = None
            for line in p.stdout.splitlines():
                if line.startswith("RESULT "):
                    res = json.loads(line[7:])
            if res is None:
                res = dict(ok=False, rc=p.returncode, stderr=p.stderr[-1500:], stdout=p.stdout[-500:])
        except subprocess.TimeoutExpired:
            res = dict(ok=False, timeout=True)
        res["secs"] = round(time.time() - t0, 1)
        report[name] = res
        print(("PASS " if res.get("ok") else "FAIL ") + name + " " + json.dumps(res)[:600], flush=True)
        with open(os.path.join(ROOT, "gpurun_out", os.environ.get("UB_PROBE_OUT", "probe.json")), "w") as f:
            json.dump(report, f, indent=1)
    nfail = sum(1 for r in report.values() if not r.get("ok"))
    print(f"{len(report) - nfail}/{len(report)} cases passed")
    sys.exit(1 if nfail else 0)
