"""Per-CUDA-source-line instruction attribution from an ncu report (needs -lineinfo + --import-source on).
usage: python profiles/ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys
rep, kn = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', 'regex:' + kn],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == 'Line No')
ie, it, isamp = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
lines = [r for r in rows if len(r) > it and r[0].isdigit() and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in lines); tots = sum(int(r[isamp]) for r in lines if r[isamp].isdigit())
print(f"== {kn}: {tot/1e6:.1f} M warp instructions, {tots} samples")
print(f"{'line':>5} {'Minst':>8} {'%inst':>6} {'%smpl':>6} {'thr/inst':>8}  source")
for r in sorted(lines, key=lambda r: -int(r[ie]))[:top]:
    n = int(r[ie])
    print(f"{r[0]:>5} {n/1e6:8.2f} {100*n/tot:6.1f} {100*(int(r[isamp]) if r[isamp].isdigit() else 0)/max(1,tots):6.1f} {int(r[it])/max(1,n):8.1f}  {r[1].strip()[:110]}")
