"""SASS excerpts of the two hot kernels from the built libtgnh.so -> profiles/sass_v2_{a,b}_<round>.txt
(cuobjdump -sass; what proves the TMA bulk copies, mbarriers and warp shuffles are in the shipped binary)."""
import collections, os, re, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "openmm_drudenose_b200", "libtgnh.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur and re.search(r"/\*[0-9a-f]{4,5}\*/", line):
        funcs[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).rstrip())
want = {"a": ("_ZN4tgnh14tgnh_v2_kernelILi0ELi0ELb1ELb1EEEvNS_10StreamArgsE", "first half: tgnh_v2_kernel<V2_A, f32 forces, COM group, hard wall>"),
        "b": ("_ZN4tgnh14tgnh_v2_kernelILi1ELi0ELb1ELb0EEEvNS_10StreamArgsE", "second half: tgnh_v2_kernel<V2_B, f32 forces, COM group>")}
keep = re.compile(r"UBLKCP|SYNCS|SHFL|STG\.E\.128|LDS\.128|ELECT|ACQBULK|MUFU|F2F|DADD|DFMA|DMUL")
for tag, (name, title) in want.items():
    ins = funcs[name]
    ops = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", l.split("*/", 1)[1].strip()).split()[0].split(".")[0] for l in ins)
    out = [f"# SASS excerpt (cuobjdump -sass, sm_100a cubin inside openmm_drudenose_b200/libtgnh.so), round {R}", f"# {title}", "",
           f"instructions: {len(ins)}; opcode histogram (top 25): " + ", ".join(f"{k} {v}" for k, v in ops.most_common(25)), "",
           "TMA bulk copies (UBLKCP), mbarrier operations (SYNCS.*), warp shuffles, 128-bit shared loads and global stores, and every XU / fp64 instruction:", ""]
    out += [l for l in ins if keep.search(l)]
    open(os.path.join(ROOT, "profiles", f"sass_v2_{tag}_{R}.txt"), "w").write("\n".join(out) + "\n")
    print(tag, len(ins), "instructions;", sum(1 for l in ins if "UBLKCP" in l), "UBLKCP,", sum(1 for l in ins if "SYNCS" in l), "SYNCS,", sum(1 for l in ins if "SHFL" in l), "SHFL")
