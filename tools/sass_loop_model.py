"""For every pair_sum_kernel instance in an object file: find the hottest loop (the backward branch whose body holds the
most DFMAs) and print the modelled FP64-pipe cycles per pair (see tools/sass_reuse_stats.py for the model)."""
import re
import subprocess
import sys


def model(lines):
    cache = [None, None, None]
    n = cyc = other = 0
    for ln in lines:
        m = re.search(r"/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?(\S+)\s+(.*?);", ln)
        if not m:
            continue
        op, args = m.group(2), m.group(3)
        if not op.startswith(("DFMA", "DMUL", "DADD")):
            other += 1
            continue
        ops = [a.strip() for a in args.split(",")][1:]
        new = [None] * 3
        need = set()  # distinct registers: the same register in two slots is read once (tools/const_probe.cu: dup probes)
        for slot, a in enumerate(ops[:3]):
            reg = re.match(r"[-|]*?(R\d+)(\.reuse)?", a)
            if not reg:
                continue
            if cache[slot] != reg.group(1):
                need.add(reg.group(1))
            if reg.group(2):
                new[slot] = reg.group(1)
        fresh = len(need)
        cache = new
        n += 1
        cyc += max(2, fresh)
    return n, cyc, other


def main():
    obj = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else "PairCfg"
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    for f in funcs:
        name = f.split("\n", 1)[0]
        if pat not in name:
            continue
        dem = subprocess.run(["c++filt", name.strip()], capture_output=True, text=True).stdout.strip()
        lines = [ln for ln in f.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/\s", ln) and not re.match(r"\s*/\* 0x", ln)]
        addr = {}
        for i, ln in enumerate(lines):
            m = re.search(r"/\*([0-9a-f]{4})\*/", ln)
            addr[int(m.group(1), 16)] = i
        best = None
        for i, ln in enumerate(lines):
            m = re.search(r"BRA(\.U)?\s+(!?U?P\d+,\s+)?0x([0-9a-f]+)", ln)
            if not m:
                continue
            tgt = int(m.group(3), 16)
            if tgt in addr and addr[tgt] < i:
                body = lines[addr[tgt]:i + 1]
                nd = sum("DFMA" in b for b in body)
                if best is None or nd > best[0]:
                    best = (nd, body)
        if best:
            n, cyc, other = model(best[1])
            mufu = sum("MUFU.RCP64H" in b for b in best[1])
            print(f"{dem[:70]:70s} fp64 {n:4d} other {other:3d} pairs {mufu:3d}  modelled {cyc / max(mufu, 1):6.2f} cyc/pair "
                  f"({n / max(mufu, 1):.1f} instr/pair)")


if __name__ == "__main__":
    main()
