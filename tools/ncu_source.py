"""Per-instruction view of one kernel in an .ncu-rep (authoring container; ncu -i needs no GPU):
    python tools/ncu_source.py REP KERNEL_REGEX [min_samples]
prints total warp instructions / samples and the hottest SASS lines (samples, executions, shared-memory conflicts)."""
import csv
import io
import subprocess
import sys


def sections(rep, regex):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{regex}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            secs.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    return secs


def main():
    rep, regex = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    for sec in sections(rep, regex)[:1]:
        ix = {h: i for i, h in enumerate(sec["hdr"])}
        data = sec["rows"]
        I = lambda r, k: int(r[ix[k]] or 0)
        print(sec["name"][:100])
        print("warp instructions", sum(I(r, "Instructions Executed") for r in data), "samples", sum(I(r, "# Samples") for r in data),
              "smem conflict wavefronts", sum(I(r, "L1 Wavefronts Shared Excessive") for r in data))
        top = sorted(range(len(data)), key=lambda k: -I(data[k], "# Samples"))[:top_n]
        for k in sorted(top):
            r = data[k]
            print(f"{k:5d} smp {I(r, '# Samples'):6d} exe {I(r, 'Instructions Executed'):9d} exc {I(r, 'L1 Wavefronts Shared Excessive'):8d}  {r[ix['Source']].strip()[:90]}")


if __name__ == "__main__":
    main()
