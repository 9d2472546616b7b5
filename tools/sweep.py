"""Residency sweep for the decode kernel: runs bench.py (device-timed leg only) under different
LZB_GROUP / LZB_UNITS settings and prints one line per setting.  Used for profiles/*_residency.md.

    python tools/sweep.py [--workload lzma64k] 32:32 16:32 16:24 8:32
"""
import json
import os
import subprocess
import sys


def main():
    args = sys.argv[1:]
    workload = "lzma64k"
    if args and args[0] == "--workload":
        workload = args[1]
        args = args[2:]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for spec in args:
        g, u = spec.split(":")
        env = dict(os.environ, LZB_GROUP=g, LZB_UNITS=u)
        p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", workload, "--no-e2e",
                            "--no-cpu", "--steps", "3", "--warmup", "3"], env=env, capture_output=True, text=True)
        line = [l for l in p.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(f"G={g} units={u}: FAILED rc={p.returncode} {p.stderr[-300:]}")
            continue
        j = json.loads(line[-1])
        print(f"{workload} G={g} units/SM={u}: {j['value']:.4f} {j['unit']} verified={j.get("verified")} "
              f"clock={j['clocks'].get('sm_mhz')}", flush=True)


if __name__ == "__main__":
    main()
