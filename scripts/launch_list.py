"""Compact an `ncu --metrics gpu__time_duration.sum --csv` log into idx,kernel,grid,block,ns (+ per-kernel shares)."""
import csv, re, sys
from collections import OrderedDict

def short(name):
    if "at::" in name or "at_cuda" in name:
        return "torch:random" if "random" in name or "distribution" in name else "torch:elementwise/other"
    m = re.match(r"void\s+(?:lg::)?(\w+)(<[^(]*>)?\(", name)
    return ("lg::" + m.group(1) + (m.group(2) or "")) if m else name[:60]

def main(path, header):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k, g, b, v = (hdr.index(x) for x in ("Kernel Name", "Grid Size", "Block Size", "Metric Value"))
    out, tot = [], OrderedDict()
    for i, r in enumerate(rows[1:]):
        name, ns = short(r[k]), int(float(r[v].replace(",", "")))
        out.append((i, name, r[g], r[b], ns))
        tot[name] = tot.get(name, 0) + ns
    print(f"# {header}")
    print("# per-launch times are cold-cache and serialised: compare SHARES")
    lg_total = sum(ns for n, ns in tot.items() if n.startswith("lg::"))
    for n, ns in tot.items():
        if n.startswith("lg::"):
            print(f"# share {n}: {100.0 * ns / lg_total:.1f}% ({ns / 1e6:.2f} ms over the capture)")
    print("idx,kernel,grid,block,ns")
    for i, n, gr, bl, ns in out:
        print(f'{i},{n},"{gr}","{bl}",{ns}')

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu launch list")
