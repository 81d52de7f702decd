#!/usr/bin/env python3
"""Fingerprint of the scoring kernels' inner loops in a built liboswald_cuda.so.

ptxas's schedule of the row sweep is sensitive to small changes elsewhere in the kernel (a
loop-carried flag in the chunk fetch cost 0.8 % at config 2, a few more instructions 1.7 %), and
box-to-box spread hides steps of that size.  This tool extracts, per kernel instance, the opcode
sequence of the innermost large loop (the step loop), the register count and the stack size;
tests/test_host.py compares them with tests/golden/kernel_schedule.json so that a change of the
schedule is noticed on the CPU and followed by an A/B run on one box (tools/gpu_ab.sh).

usage: kernel_schedule.py [lib.so]            print the fingerprints as JSON
       kernel_schedule.py --update [lib.so]   rewrite tests/golden/kernel_schedule.json
"""
import collections
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "kernel_schedule.json")
# the instances the headline workload and the single-query path run at full size
KERNELS = {
    "two_track_G32_R40": "_ZN7osw_u1613sw_u16_kernelILi32ELi40ELi384ELb0ELb0ELb0EEEvNS_5KArgsE",
    "two_track_G32_R20": "_ZN7osw_u1613sw_u16_kernelILi32ELi20ELi512ELb0ELb0ELb0EEEvNS_5KArgsE",
    "pair_db_G32_R28": "_ZN7osw_u1613sw_u16_kernelILi32ELi28ELi512ELb1ELb0ELb0EEEvNS_5KArgsE",
    "pair_db_G4_R36": "_ZN7osw_u1613sw_u16_kernelILi4ELi36ELi384ELb1ELb0ELb0EEEvNS_5KArgsE",
    # the transposed form's 16-warp instance: the sweep of 4 rows per lane that hands its bottom row on in
    # place, both ends of the array inside the query (two steps per trip: 24 add-max, one ring read and one
    # ring write per step) - what small databases spend their time in
    "transposed_R4_in_place": ("_ZN7osw_t1613sw_t16_kernelILi4EEEvNS_5TArgsE", {"VIADDMNMX.U16x2": 24, "STS.64": 2, "LDS.64": 2}),
}


def step_loop(lib, fun, select=None):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, lib], capture_output=True, text=True, check=True).stdout
    ops = []
    for line in out.splitlines():
        m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ops.append((int(m.group(1), 16), m.group(2)))
    loops = []
    for addr, text in ops:
        if "BRA" in text:
            m = re.search(r"0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) < addr:
                loops.append((int(m.group(1), 16), addr))
    # the step loop: the smallest loop that holds the row sweep (a dozen or more DPX add-max instructions)
    def dpx(lo, hi):
        return sum(1 for a, t in ops if lo <= a <= hi and "VIADDMNMX" in t)
    big = sorted((hi - lo, lo, hi) for lo, hi in loops if dpx(lo, hi) >= 12)
    if select:        # ... or the smallest loop with exactly these opcode counts
        def mix(lo, hi):
            return collections.Counter((t.split()[1] if t.startswith("@") else t.split()[0]) for a, t in ops if lo <= a <= hi)
        big = [(n, lo, hi) for n, lo, hi in big if all(mix(lo, hi).get(k, 0) == v for k, v in select.items())]
    if not big:
        raise RuntimeError("no step loop found in " + fun)
    _, lo, hi = big[0]
    return [(t.split()[1] if t.startswith("@") else t.split()[0]) for a, t in ops if lo <= a <= hi]


def resources(lib):
    out = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True, check=True).stdout
    res, fun = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fun = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and fun:
            res[fun] = (int(m.group(1)), int(m.group(2)))
    return res


def fingerprints(lib):
    res = resources(lib)
    out = {}
    for name, fun in KERNELS.items():
        select = None
        if isinstance(fun, tuple):
            fun, select = fun
        seq = step_loop(lib, fun, select)
        mix = collections.Counter(seq)
        out[name] = {"registers": res[fun][0], "stack": res[fun][1], "loop_instructions": len(seq),
                     "viaddmnmx_u16x2": mix.get("VIADDMNMX.U16x2", 0), "vimnmx3_u16x2": mix.get("VIMNMX3.U16x2", 0),
                     "lds128": mix.get("LDS.128", 0), "local_memory_ops": sum(v for k, v in mix.items() if k.startswith(("LDL", "STL"))),
                     "schedule_sha1": hashlib.sha1(" ".join(seq).encode()).hexdigest()}
    return out


def main():
    args = [a for a in sys.argv[1:] if a != "--update"]
    lib = args[0] if args else os.path.join(ROOT, "oswald_b200", "liboswald_cuda.so")
    fp = fingerprints(lib)
    if "--update" in sys.argv:
        json.dump(fp, open(GOLDEN, "w"), indent=1, sort_keys=True)
        print("wrote", GOLDEN)
    else:
        print(json.dumps(fp, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
