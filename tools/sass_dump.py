"""SASS listing of one kernel of libpemspgemm.so (cuobjdump -sass) with a mnemonic histogram on top.
    python tools/sass_dump.py '<demangled-name regex>' profiles/r02_sass_<name>.txt"""
import collections, re, subprocess, sys
LIB = "pem_spgemm_b200/lib/libpemspgemm.so"
rx, out = re.compile(sys.argv[1]), sys.argv[2]
names = subprocess.run(f"cuobjdump -sass {LIB} | grep 'Function :'", shell=True, capture_output=True, text=True).stdout.split("\n")
mangled = [n.split("Function : ")[1].strip() for n in names if "Function : " in n]
dem = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True).stdout.split("\n")
hits = [(m, d) for m, d in zip(mangled, dem) if rx.search(d)]
if len(hits) != 1:
    sys.exit(f"{len(hits)} kernels match: " + "; ".join(d[:80] for _, d in hits))
m, d = hits[0]
sass = subprocess.run(["cuobjdump", "-sass", "-fun", m, LIB], capture_output=True, text=True).stdout
lines = [re.sub(r"/\* 0x[0-9a-f]{16} \*/", "", l).rstrip() for l in sass.split("\n")]
lines = [l for l in lines if re.match(r"\s+/\*[0-9a-f]{4}\*/", l) or l.strip().startswith(".L_")]
ops = collections.Counter(re.sub(r"@!?U?P\d\s+", "", l.split("*/", 1)[1].strip()).split()[0].rstrip(";") for l in lines if "*/" in l)
with open(out, "w") as f:
    f.write(f"# {d}\n# cuobjdump -sass of {LIB} (sm_100a), {sum(ops.values())} instructions\n")
    f.write("# mnemonics: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(28)) + "\n")
    f.write("\n".join(lines) + "\n")
print(out, sum(ops.values()), "instructions;", ", ".join(f"{k} {v}" for k, v in ops.most_common(10)))
