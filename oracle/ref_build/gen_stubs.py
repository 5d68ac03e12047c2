"""gen_stubs.py LIB OUT.s [EXTRA_LIB...] — assembly source defining, as weak trapping functions, every symbol LIB still
leaves undefined that neither the system libraries nor EXTRA_LIBs provide (symbols of reference translation units
the hot path never calls: diagnostics, I/O, MPI wrappers, Python).  A call into one prints its name and aborts."""
import glob
import subprocess
import sys

lib, out, extra = sys.argv[1], sys.argv[2], sys.argv[3:]
und = subprocess.run(["nm", "-D", "--undefined-only", lib], capture_output=True, text=True).stdout.split("\n")
syms = [l.split()[-1] for l in und if l.strip() and l.split()[-2] != "w"]  # weak undefined stay undefined
provided = set()
libs = list(extra)
for pat in ("/usr/lib/x86_64-linux-gnu/libstdc++.so.6", "/lib/x86_64-linux-gnu/libc.so.6", "/lib/x86_64-linux-gnu/libm.so.6",
            "/usr/lib/x86_64-linux-gnu/libgomp.so.1", "/lib/x86_64-linux-gnu/libgcc_s.so.1"):
    libs += glob.glob(pat)
for l in libs:
    o = subprocess.run(["nm", "-D", "--defined-only", l], capture_output=True, text=True).stdout
    for line in o.split("\n"):
        p = line.split()
        if p:
            provided.add(p[-1].split("@")[0])
need = [s for s in syms if s.split("@")[0] not in provided and not s.startswith("Py") and not s.startswith("_Py")]
py = [s for s in syms if s.startswith("Py") or s.startswith("_Py")]
with open(out, "w") as f:
    f.write("\t.text\n")
    for i, s in enumerate(need + py):
        f.write(f"\t.weak {s}\n\t.type {s}, @function\n{s}:\n\tleaq .Lname{i}(%rip), %rdi\n\tcall sb200_ref_trap_report@PLT\n\tud2\n")
    f.write("\t.section .rodata\n")
    for i, s in enumerate(need + py):
        f.write(f".Lname{i}:\n\t.string \"{s}\"\n")
    f.write('\t.section .note.GNU-stack,"",@progbits\n')
print(f"gen_stubs: {len(need)+len(py)} unresolved symbols routed to the trap stub")
