"""Dynamic opcode mix of a kernel from an ncu report's source page (needs --import-source on / -lineinfo not required for SASS view).
usage: python tools/ncu_opmix.py report.ncu-rep [units]   -> thread-level instruction counts per opcode class, per `units` (e.g. permutations)"""
import collections, csv, io, re, subprocess, sys

def opmix(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # find header
    hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Instructions Executed" in c for c in r))
    hdr = rows[hi]
    si = hdr.index("Source")
    ei = next(i for i, c in enumerate(hdr) if c.strip() == "# Instructions Executed" or c.strip() == "Instructions Executed")
    mix = collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= max(si, ei):
            continue
        m = re.match(r"\s*(?:@!?U?P\d\s+)?([A-Z0-9_.]+)(.*)", r[si])
        if not m:
            continue
        try:
            n = int(float(r[ei]))
        except ValueError:
            continue
        op = m.group(1)
        rest = m.group(2)
        if op.startswith("IMAD.WIDE"):
            form = "rz" if rest.strip().rstrip(";").strip().endswith("RZ") else "add"
            if ".X" in op: form = "x"
            elif re.search(r"R\d+, P\d", rest): form += "+cc"
            if re.search(r", (0x[0-9a-f]+|-0x[0-9a-f]+),", rest): form += "+imm"
            op = "IMAD.WIDE[" + form + "]"
        mix[op] += n
    return mix

if __name__ == "__main__":
    mix = opmix(sys.argv[1])
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    tot = sum(mix.values())
    for op, n in mix.most_common(40):
        print("%-28s %14.1f  %5.1f%%" % (op, n * 32 / units, 100.0 * n / tot))
    print("%-28s %14.1f" % ("total (thread-instr/unit)", tot * 32 / units))
