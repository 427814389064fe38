"""Count the SASS instructions that show what each kernel of libvggp.so is built from (run here, no GPU needed):

    python tools/sass_markers.py > profiles/r2_sass_markers.txt

Per kernel: instruction count and the number of bulk-copy (TMA) instructions UBLKCP, mbarrier operations SYNCS, FP64 tensor
instructions DMMA, tcgen05 tensor instructions UTC*MMA / tensor-memory accesses LDTM / STTM, tensor-map TMA UTMALDG / UTMASTG,
16-byte global loads LDG.E.128, fire-and-forget reductions REDG, multimem in-switch reductions LDGMC
(multimem.ld_reduce; multimem.st is an ordinary STG to the multicast address), shuffles, cp.async LDGSTS, and the FP64 / FP32 FMA and
special-function counts."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "variational-gridded-gaussian-processes_b200", "libvggp.so")
MARK = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("DMMA", r"\bDMMA"), ("UTCMMA", r"\bUTC\w*MMA"), ("LDTM/STTM", r"\b(LDTM|STTM)"),
        ("UTMALDG/STG", r"\bUTMA(LDG|STG)"), ("LDG.128", r"\bLDG\.E\.128|\bLDG\.E\.\w*\.?128|LDG\.E\.EF\.128|LDG.*\.128"),
        ("REDG", r"\bREDG?\."), ("LDGMC", r"\bLDGMC"), ("SHFL", r"\bSHFL"), ("LDGSTS", r"\bLDGSTS"),
        ("DFMA", r"\bDFMA"), ("FFMA", r"\bFFMA"), ("MUFU", r"\bMUFU")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(1)
        kernels[cur]["n"] += 1
        for name, pat in MARK:
            if re.search(pat, ins):
                kernels[cur][name] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}   architectures in the fat binary: {', '.join(arch)}")
    print("# columns: instructions | " + " | ".join(n for n, _ in MARK))
    tot = collections.Counter()
    for (k, c), d in zip(kernels.items(), demangled):
        d = re.sub(r"^void ", "", d)
        d = re.sub(r"\(.*", "", d)
        print(f"{d[:64]:64s} {c['n']:6d} | " + " | ".join(f"{c[n]:5d}" for n, _ in MARK))
        tot.update(c)
    print(f"{'TOTAL (' + str(len(kernels)) + ' kernels)':64s} {tot['n']:6d} | " + " | ".join(f"{tot[n]:5d}" for n, _ in MARK))


if __name__ == "__main__":
    main()
