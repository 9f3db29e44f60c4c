#!/usr/bin/env python
"""SASS evidence for profiles/: per-kernel instruction counts of the mnemonics that prove the Blackwell paths
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc, UBLKCP = TMA bulk copy,
SYNCS = mbarrier, FFMA2 / FADD2 = packed fp32, REDUX, MUFU.*) plus an excerpt around the first UTCHMMA of each kernel.
usage: sass_counts.py <libmsort.so> > profiles/sass_r02_kernels.txt   (needs cuobjdump)"""
import collections, re, subprocess, sys

WANT = ["UTCHMMA", "LDTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FFMA", "HFMA2", "REDUX", "MUFU.TANH", "MUFU.EX2",
        "MUFU.RCP", "MUFU.LG2", "IMAD", "DFMA", "DADD", "DMUL", "LDG", "STG", "LDS", "STS", "BAR", "STL", "LDL"]
KEEP = ("step_kernelILi3ELi0ELi2ELb1ELb1ELb1ELb0ELb0ELb0E", "step_kernelILi3ELi0ELi2ELb1ELb1ELb1ELb0ELb1ELb0E", "step_kernelILi2ELi0ELi2ELb1ELb1ELb1ELb1ELb0ELb0E",
        "step_kernelILi2ELi0ELi2ELb1ELb1ELb1ELb0ELb0ELb0E", "step_kernelILi1ELi0ELi2ELb1ELb1ELb1ELb0ELb0ELb0E", "rollout_policy_kernel", "policy_act_kernelILi29ELi22E",
        "ppo_kernelILi29ELi22ELb1E", "ppo_kernelILi29ELi22ELb0E", "pack_fused_kernel", "press_policy_kernel", "step_kernelILi2ELi0ELi2ELb1ELb1ELb1ELb0ELb0ELb1E", "tc_logits_kernel", "adam_kernel", "gae_kernel")

def main():
    out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout.split("\n")
    starts = [i for i, l in enumerate(out) if "Function :" in l] + [len(out)]
    print(f"# cuobjdump -sass {sys.argv[1].split('/')[-1]} (sm_100a): instruction counts per kernel\n")
    for a, b in zip(starts[:-1], starts[1:]):
        name = out[a].split("Function :")[1].strip()
        if not any(k in name for k in KEEP):
            continue
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:150]
        ins = [re.sub(r"^\s*/\*[0-9a-f]+\*/\s+", "", l).split("/*")[0].strip() for l in out[a:b] if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]
        c = collections.Counter()
        for x in ins:
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", x)
            if not m:
                continue
            op = m.group(2)
            for w in WANT:
                if op == w or op.startswith(w + ".") or (w.startswith("MUFU") and op.startswith(w)):
                    c[w] += 1
        print(f"## {demangled}\n   {len(ins)} instructions; " + ", ".join(f"{w} {c[w]}" for w in WANT if c[w]))
        idx = [k for k, x in enumerate(ins) if "UTCHMMA" in x]
        if idx:
            k = idx[0]
            print("   first tcgen05.mma and its neighbourhood:")
            for x in ins[max(0, k - 6):k + 8]:
                print("      " + x)
        print()

if __name__ == "__main__":
    main()
