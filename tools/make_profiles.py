#!/usr/bin/env python
"""profiles/r02_*.txt from the scratch captures of tools/gpu_profiles.sh (gpurun_out/prof_<tag>_<kernel>.ncu-rep, <tag>_launches.csv).
usage: make_profiles.py <tag> [round label]      run in the authoring container (needs ncu, cuobjdump, nvdisasm and the built objects)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
label = sys.argv[2] if len(sys.argv) > 2 else "r02"
KERNELS = [("score_dense_kernel", "score.o", "score_dense_kernelILb1ELi6"), ("lookup_kernel", "lookup.o", "lookup_kernelILb0ELb1ELb1"),
           ("extract_identify_kernel", "extract.o", "extract_identify_kernelILb0ELb1")]


def run(cmd):
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True).stdout


for name, obj, mangled in KERNELS:
    rep = os.path.join("gpurun_out", "prof_%s_%s.ncu-rep" % (tag, name))
    if not os.path.exists(os.path.join(ROOT, rep)):
        print("missing", rep)
        continue
    out = ("# ncu --set full --clock-control none --import-source on -k regex:%s --launch-skip 3 --launch-count 1 (inside the timed identify step)\n"
           "#   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e   (the bench's FULL C2 size: 6.7 M reads = 1.742 G k-mers vs 1.67 G entries)\n"
           "#   round 2, final kernels; .ncu-rep: %s (scratch, not tracked); written by tools/make_profiles.py\n" % (name, rep))
    out += run([sys.executable, "tools/ncu_summary.py", rep])
    out += "\n# hottest source lines (share of warp-state samples / of executed warp instructions, avg active threads)\n"
    out += run([sys.executable, "tools/ncu_lines.py", rep, os.path.join("kasa_b200", "csrc", "build", obj), mangled, "24"])
    open(os.path.join(ROOT, "profiles", "%s_ncu_%s.txt" % (label, name)), "w").write(out)
    print("wrote", name)
csv = os.path.join("gpurun_out", "%s_launches.csv" % tag)
if os.path.exists(os.path.join(ROOT, csv)):
    out = ("# ncu launch list (gpu__time_duration.sum, --clock-control none), round 2, final kernels, of:\n"
           "#   ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e\n"
           "#   (C2: 50 taxa x 33.4 Mbp index = 1.67 G entries, 6.7 M reads = 1.742 G k-mers per step; 5 identify steps captured)\n"
           "# includes the untimed set-up (index build: extract_build / hist / 12 onesweep passes / dedup / trie / group masks; read generation).\n"
           "# per-launch times are cold-cache and serialised: compare SHARES with bench.py's CUDA-event stage times (DESIGN.md §4).\n")
    out += run([sys.executable, "tools/launch_summary.py", csv])
    open(os.path.join(ROOT, "profiles", "%s_launches_bench_full.txt" % label), "w").write(out)
    print("wrote launch list")
