"""Per-source-line summary of an ncu report (needs -lineinfo + --import-source on): executed warp instructions and
stall samples per CUDA line, for the kernels of one launch.   python tools/prof_lines.py rep.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; out = []
for r in rows:
    if r and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 7 and r[2] == '-' and r[6].isdigit():
        out.append((int(r[6]), int(r[7] or 0), cur, r[0], r[1]))
tot = sum(o[0] for o in out) or 1; tote = sum(o[1] for o in out) or 1
print('total samples', tot, 'total executed warp instructions', tote)
for s, e, f, ln, src in sorted(out, key=lambda o: -o[1])[:top]:
    print(f'{100*s/tot:5.1f}%smp {100*e/tote:5.1f}%exe {e:>10} {f}:{ln:>5} {src.strip()[:100]}')
