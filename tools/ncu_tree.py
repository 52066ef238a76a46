#!/usr/bin/env python3
"""The reference's timing tree (src/DistributedMatrixVector.chpl:1028-1052, src/BatchedOperator.chpl:53-57) for a FUSED
kernel: the stages cannot be timed with events, so the warp instructions and the stall samples of an ncu capture
(--import-source on, -lineinfo) are attributed to the reference's timers through the device function every source line
belongs to:

    computeOffDiag.applyOffDiag   term kernel: emit test, pop, coefficient       (row_terms, pop_term, bp_gather, ...)
    computeOffDiag.stateInfo      orbit minimum / canonical forms                (orbit_*, min_rotation_*, butterfly, ...)
    computeOffDiag.localeIdxOf    hash64_01 % P, destination routing             (hash64_01, locale_idx_of, route)
    localProcess.indexing         state -> index / slot                          (locate*, lin_index, table_slot, bucket_load, ...)
    localProcess.accessing        y += c x: FMA, atomics, vector loads           (axpy, fma_to, atomic_accumulate, load_x, ...)
    localDiagonal                 diagonal terms                                 (diagonal)
    kernel body                   everything else (loop control, pipeline shifts, staging of the tables)

Usage: python tools/ncu_tree.py REPORT.ncu-rep  ->  markdown table (share of instructions, share of stall samples)."""
import csv
import io
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "distributed_matvec_b200", "csrc")
STAGES = [
    ("computeOffDiag.applyOffDiag", r"row_terms|pop_term|bp_gather|generic_coefficient|lut_index"),
    ("computeOffDiag.stateInfo", r"orbit_|min_rotation|butterfly|rotl_n|reverse_bits_n|top_bit|low_bit|translation_canon"),
    ("computeOffDiag.localeIdxOf", r"hash64_01|locale_idx_of|^route$"),
    ("localProcess.indexing", r"locate|lin_index|table_slot|bucket_load|slot_load|combinadic"),
    ("localProcess.accessing", r"axpy|fma_to|atomic_accumulate|smem_add|load_x|ldx|v_mul|v_scale|v_add|finish"),
    ("localDiagonal", r"^diagonal$"),
]
FUNC = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:__host__\s+|__device__\s+|__global__\s+|__forceinline__\s+|__noinline__\s+|static\s+|inline\s+)*"
                  r"[\w:<>\*&,\s]+?\b(\w+)\s*\([^;]*$")


def function_of_lines(path):
    """line number -> name of the enclosing top-level function (brace counting; good enough for this code base)."""
    out, depth, current, pending = {}, 0, None, None
    with open(path) as f:
        for no, line in enumerate(f, 1):
            code = line.split("//")[0]
            if current is None:
                m = FUNC.match(code)
                if m and m.group(1) not in ("if", "for", "while", "switch", "return", "sizeof", "asm"):
                    pending = m.group(1)
            opens, closes = code.count("{"), code.count("}")
            if pending and opens and current is None:
                current, start_depth = pending, depth
                pending = None
            out[no] = current
            depth += opens - closes
            if current is not None and depth <= start_depth:
                current = None
            if ";" in code and current is None and not opens:
                pending = None
    return out


def main():
    rep = sys.argv[1]
    text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                          capture_output=True, text=True).stdout
    funcs = {}
    inst, samp = defaultdict(int), defaultdict(int)
    cur_file = cur_line = None
    header = None
    for r in csv.reader(io.StringIO(text)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            header = r
            i_inst, i_samp = header.index("Instructions Executed"), header.index("# Samples")
            continue
        if header is None or len(r) < len(header):
            continue
        if r[0] != "":
            cur_line = int(r[0])
            continue
        try:
            a, b = int(r[i_inst]), int(r[i_samp])
        except ValueError:
            continue
        if cur_file not in funcs:
            path = os.path.join(CSRC, cur_file)
            funcs[cur_file] = function_of_lines(path) if os.path.exists(path) else {}
        fn = funcs[cur_file].get(cur_line) or ""
        stage = "kernel body"
        for name, pat in STAGES:
            if re.search(pat, fn):
                stage = name
                break
        if not fn and not funcs[cur_file]:
            stage = "toolkit headers (intrinsics)"
        inst[stage] += a
        samp[stage] += b
    ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
    print("| reference timer | warp instructions % | stall samples % |\n|---|---|---|")
    for name in [s for s, _ in STAGES] + ["kernel body", "toolkit headers (intrinsics)"]:
        if inst[name] or samp[name]:
            print(f"| {name} | {100 * inst[name] / ti:.1f} | {100 * samp[name] / ts:.1f} |")


if __name__ == "__main__":
    main()
