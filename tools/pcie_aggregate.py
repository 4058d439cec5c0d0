#!/usr/bin/env python
"""Aggregate host <-> device copy bandwidth of the box: memcpy only, N = 1, 2, 4, 8 GPUs at
the same time, pinned buffers, both directions concurrently.

Runs bin/gps_pcie_probe (tools/pcie_probe.cu) over GPU sets, directions, allocation modes and
drivers (threads of one process / one process per GPU) and writes one JSON document with the
box topology next to the numbers.  The end-to-end floor of the scan follows from it: a step
moves 4 B/key each way, so floor_Gkeys/s = min(h2d_total, d2h_total under `both`) / 4.

  python tools/pcie_aggregate.py [--out gpurun_out/pcie_aggregate.json] [--quick]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROBE = os.path.join(ROOT, "gpuprefixsums_b200", "bin", "gps_pcie_probe")


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout
    except Exception as e:  # noqa: BLE001
        return f"unavailable: {e}"


def probe(gpus, **kw):
    cmd = [PROBE, "--gpus", ",".join(map(str, gpus))]
    for k, v in kw.items():
        cmd += [f"--{k}", str(v)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    try:
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:  # noqa: BLE001
        return {"gpus": gpus, "error": (r.stderr or r.stdout)[-300:], **kw}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "pcie_aggregate.json"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    n_gpu = int(sh("nvidia-smi -L | wc -l").strip() or 1)
    doc = {
        "what": "memcpy-only probe, pinned host buffers, 256 MiB copies, ~0.6 s per point (tools/pcie_probe.cu)",
        "n_gpus_visible": n_gpu,
        "topology": sh("nvidia-smi topo -m"),
        "lscpu": sh("lscpu | egrep 'Model name|Socket|NUMA|^CPU\\(s\\)|Thread|Core|Hypervisor|L3'"),
        "numa": sh("numactl -H 2>/dev/null || cat /sys/devices/system/node/online"),
        "meminfo": sh("egrep 'MemTotal|MemAvailable|Huge|AnonHuge' /proc/meminfo"),
        "pcie_links": sh("nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.gen.max,"
                         "pcie.link.width.current --format=csv,noheader"),
        "thp": sh("cat /sys/kernel/mm/transparent_hugepage/enabled"),
        "points": [],
    }
    sets = [[0]]
    for n in (2, 4, 8):
        if n <= n_gpu:
            sets.append(list(range(n)))
    # 1) the headline matrix: N = 1, 2, 4, 8 GPUs, both directions concurrently and each alone
    for gpus in sets:
        for d in ("both", "h2d", "d2h"):
            doc["points"].append(probe(gpus, dir=d))
            print(json.dumps(doc["points"][-1]), flush=True)
    # 2) does it matter WHICH GPUs (PCIe switches / sockets)?
    if n_gpu >= 8 and not args.quick:
        for gpus in ([0, 4], [0, 2, 4, 6], [4, 5, 6, 7], [7]):
            doc["points"].append(probe(gpus, dir="both"))
            print(json.dumps(doc["points"][-1]), flush=True)
    # 3) what might move the ceiling, on the largest set: allocation mode, CPU binding, driver
    big = list(range(min(n_gpu, 8)))
    variants = [dict(alloc="huge"), dict(alloc="wc"), dict(bind=1), dict(drive="procs"), dict(drive="procs", bind=1),
                dict(mib=32)]
    if args.quick:
        variants = variants[:4]
    for v in variants:
        doc["points"].append(probe(big, dir="both", **v))
        print(json.dumps(doc["points"][-1]), flush=True)
    # summary: aggregate under `both`, default mode, per GPU count
    summ = {}
    for p in doc["points"]:
        if p.get("dir") == "both" and p.get("alloc") == "pinned" and p.get("drive") == "threads" and not p.get("bind") \
                and p.get("mib") == 256 and p.get("gpus") == list(range(len(p.get("gpus", [])))):
            n = len(p["gpus"])
            summ[str(n)] = {"h2d_total_gbs": p["h2d_total"], "d2h_total_gbs": p["d2h_total"],
                            "e2e_floor_gkeys": min(p["h2d_total"], p["d2h_total"]) / 4.0}
    doc["summary_both_directions"] = summ
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(summ))


if __name__ == "__main__":
    sys.exit(main())
