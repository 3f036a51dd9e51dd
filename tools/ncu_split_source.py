"""Split `ncu --page source --csv` output (stdin) into per-launch files; keeps only the launch indices given on the
command line: python tools/ncu_split_source.py <out_prefix> 0 3 4 ...  (each launch appears once per source view; the
first view of every launch is kept)."""
import gzip
import sys

prefix, keep = sys.argv[1], {int(v) for v in sys.argv[2:]}
idx, out = -1, None
for line in sys.stdin:
    if line.startswith('"Kernel Name"'):
        # ncu prints every launch twice (SASS view + source view); count distinct launches by pairs
        idx += 1
        if out:
            out.close()
            out = None
        launch = idx // 2
        if idx % 2 == 0 and launch in keep:
            out = gzip.open("%s_%d.csv.gz" % (prefix, launch), "wt")
    if out:
        out.write(line)
if out:
    out.close()
