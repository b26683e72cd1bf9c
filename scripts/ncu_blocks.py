"""Basic-block profile of one launch in an .ncu-rep: consecutive SASS instructions with the
same execution count are one block; prints warp-executions, average active threads, share of
the issued instructions and of the stall samples, and the opcode mix of each hot block.
usage: ncu_blocks.py REP [launch_index] [min_share_pct]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0; min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
sections = []
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Kernel Name": sections.append({"name": r[1], "rows": []}); continue
    if r and r[0] == "Address": sections[-1]["hdr"] = r; continue
    if sections and r: sections[-1]["rows"].append(r)
sec = sections[launch]
name, h, body = sec["name"], sec["hdr"], sec["rows"]
iI, iT, iS, iSrc = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples"), h.index("Source")
tot = sum(int(r[iI]) for r in body); tot_t = sum(int(r[iT]) for r in body); tot_s = max(1, sum(int(r[iS]) for r in body))
print(f"{name[:90]}\nwarp-inst {tot}  thread-inst {tot_t}  avg threads {tot_t / tot:.2f}  samples {tot_s}")
blocks, cur = [], None
for k, r in enumerate(body):
    n = int(r[iI])
    if cur is None or cur["n"] != n:
        cur = {"first": k, "n": n, "rows": []}; blocks.append(cur)
    cur["rows"].append(r)
print(f"{'sass#':>6} {'len':>4} {'execs':>10} {'thr':>5} {'%inst':>6} {'%smp':>6}  ops")
for b in blocks:
    wi = sum(int(r[iI]) for r in b["rows"]); ti = sum(int(r[iT]) for r in b["rows"]); sm = sum(int(r[iS]) for r in b["rows"])
    if wi / tot * 100 < min_share and sm / tot_s * 100 < min_share: continue
    ops = collections.Counter()
    for r in b["rows"]:
        toks = r[iSrc].split(); op = toks[1] if toks[0].startswith("@") else toks[0]; ops[op.split(".")[0]] += 1
    print(f"{b['first']:6d} {len(b['rows']):4d} {b['n']:10d} {ti / max(wi, 1):5.1f} {wi / tot * 100:6.2f} {sm / tot_s * 100:6.2f}  " +
          " ".join(f"{o}{c}" for o, c in ops.most_common(8)))
