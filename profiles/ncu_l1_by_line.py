import collections, csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass"], capture_output=True, text=True).stdout
cur = fn = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, ""])
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 20 and r[2] == "-" and r[0].isdigit() and kern in (fn or ""):
        g = r[hdr.index("L1 Tag Requests Global")]; s = r[hdr.index("L1 Wavefronts Shared")]
        a = agg[(cur, r[0])]
        a[0] += int(g or 0); a[1] += int(s or 0); a[2] = r[1][:100]
tg = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("total L1 tag requests global", tg, " L1 wavefronts shared", ts)
for k, v in sorted(agg.items(), key=lambda kv: -(kv[1][0] + kv[1][1]))[:28]:
    print(f"  glob {v[0]:9d} ({100*v[0]/max(tg,1):4.1f}%)  shared {v[1]:9d} ({100*v[1]/max(ts,1):4.1f}%)  {k[0]}:{k[1]}  {v[2]}")
