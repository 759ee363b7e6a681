"""profiles/traffic.json from an ncu --set full capture of the bench command: dram__bytes_read.sum + dram__bytes_write.sum
per launch of every kernel (averaged over the captured launches), with the hash of the kernel source the capture was
taken from, so that bench.py reports `roofline.traffic` only while that source is unchanged."""
import csv, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for r in rows[2:]:
    name = r[ik].split("(")[0]
    v = float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]
    acc.setdefault(name, []).append(v)
res = {k: round(sum(v) / len(v)) for k, v in acc.items()}
src = os.path.join(ROOT, "zlib_b200", "csrc", "zb_deflate.cu")
res["_source_sha16"] = hashlib.sha256(open(src, "rb").read()).hexdigest()[:16]
res["_note"] = note or f"dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the launches in {os.path.basename(rep)}"
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
