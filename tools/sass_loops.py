"""Static SASS view of one kernel: backward branches (loops) with their instruction counts and opcode mix.
usage: python tools/sass_loops.py file.o|.so kernel-substring [min_len]"""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]; minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
cur = None; funcs = {}
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m: cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if m and cur: funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name: continue
    print(name, len(ins), "instructions")
    addr = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, s) in enumerate(ins):
        m = re.search(r'BRA(?:\.\w+)*\s+(?:!?U?P\d,\s*)?(?:`\(\.L_x_\d+\)|0x([0-9a-f]+))', s)
        if not m or not m.group(1): continue
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr and i - addr[tgt] >= minlen:
            body = ins[addr[tgt]:i + 1]
            ops = collections.Counter((x.split()[1] if x.startswith('@') else x.split()[0]).split('.')[0] for _, x in body)
            print(f"  loop {tgt:#x}..{a:#x}: {len(body)} instr; " + " ".join(f"{k}={v}" for k, v in ops.most_common(14)))
