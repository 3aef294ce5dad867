"""Per-source-line summary of an .ncu-rep captured with --import-source on: shared-memory wavefronts (ideal / excessive),
instructions and stall samples per CUDA source line of one kernel.
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [n_elements] [top]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    nel = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # the source page is a sequence of (file, function) blocks; all files of ONE launch of a function are merged, a file
    # path seen twice for the same function starts the next launch
    blocks, cur, fpath = [], None, ""
    for r in rows:
        if r and r[0] == "File Path":
            fpath = r[1].split("/")[-1] if len(r) > 1 else ""
            continue
        if r and r[0] == "Function Name":
            if cur is not None and cur["name"] == r[1] and fpath not in cur["files"]:
                cur["files"].add(fpath)
                cur["file"] = fpath
            else:
                cur = {"name": r[1], "rows": [], "hdr": None, "file": fpath, "files": {fpath}}
                blocks.append(cur)
            continue
        if cur is None:
            continue
        if r and r[0] == "Line No":
            cur["hdr"] = r
            continue
        if cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["rows"].append([cur["file"] + ":" + r[0]] + r[1:])
    done = set()
    for b in blocks:
        if not b["rows"] or b["name"] in done:
            continue
        done.add(b["name"])
        h = b["hdr"]
        iw, ie, ii, isamp = (h.index(k) for k in ("L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive",
                                                  "Instructions Executed", "# Samples"))
        stall = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
        agg = {}
        for r in b["rows"]:
            if r[2] != "-":
                continue
            try:
                agg[r[0]] = (float(r[iw] or 0), float(r[ie] or 0), float(r[ii] or 0), float(r[isamp] or 0), r[1],
                             {k: float(r[h.index(k)] or 0) for k in stall})
            except ValueError:
                pass
        tw, te, tn, ts = (sum(v[i] for v in agg.values()) for i in range(4))
        print("== %s" % b["name"][:100])
        print("per element: shared wavefronts %.0f (excessive %.0f), warp instructions %.0f, samples %d" % (
            tw / nel, te / nel, tn / nel, ts))
        st = {k: sum(v[5][k] for v in agg.values()) for k in stall}
        print("stalls: " + ", ".join("%s %.1f%%" % (k[6:], 100 * v / max(ts, 1)) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
        for l, v in sorted(agg.items(), key=lambda kv: -(kv[1][3]))[:top]:
            big = max(v[5].items(), key=lambda kv: kv[1])
            print("%26s samp %4.1f%% (%s) wf/el %6.0f exc %6.0f inst/el %6.0f | %s" % (
                l, 100 * v[3] / max(ts, 1), big[0][6:], v[0] / nel, v[1] / nel, v[2] / nel, v[4].strip()[:90]))


if __name__ == "__main__":
    main()
