"""Turn an .ncu-rep (ncu --set full) into the two files kept under profiles/:
  <out>.csv          metric, unit, one column per profiled kernel   (from --page raw --csv, transposed)
  <out>_details.txt  the 'details' page (sections, rules, estimated speed-ups)
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/ncu_x_r01"""
import csv
import io
import re
import subprocess
import sys


def short(name):
    m = re.search(r"(\w+)<([^>]*)>\(", name)
    return "%s<%s>" % (m.group(1), m.group(2).replace(" ", "").replace("(int)", "").replace("(bool)", "")) if m else name[:60]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kcol = hdr.index("Kernel Name")
    names = [short(r[kcol]) for r in data]
    with open(out + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + names)
        for c in range(len(hdr)):
            if hdr[c] in ("ID", "Process ID", "Process Name", "Host Name", "Context", "Stream", "Device", "CC"):
                continue
            w.writerow([hdr[c], units[c]] + [r[c] for r in data])
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True, check=True).stdout
    open(out + "_details.txt", "w").write(det)
    print("wrote", out + ".csv", out + "_details.txt", names)


if __name__ == "__main__":
    main()
