set +e
python - <<PY
import os, sys, numpy as np
sys.path.insert(0, ".")
from rabbittclust_b200 import synth
os.makedirs("/tmp/c1/genomes", exist_ok=True)
gen = synth.genome_families(50, 20, 1000000, seed=12345)
files = []
for i, g in enumerate(gen):
    p = "/tmp/c1/genomes/g%05d.fna" % i
    a = np.frombuffer(g, np.uint8)
    rows = a[: len(a) // 80 * 80].reshape(-1, 80)
    body = np.concatenate([rows, np.full((len(rows), 1), 10, np.uint8)], axis=1).tobytes() + bytes(a[len(a) // 80 * 80:]) + b"\n"
    open(p, "wb").write((">g%05d synthetic\n" % i).encode() + body)
    files.append(p)
open("/tmp/c1/list.txt", "w").write("\n".join(files) + "\n")
PY
cd /tmp/c1
ls genomes | wc -l
for it in 1 2; do ( time TCLUST_TIMING=1 $GRAFT_REPO_ROOT/rabbittclust_b200/host/bin/clust-mst -l -i list.txt -k 21 -s 1000 -d 0.05 -o out.txt ) > log$it.txt 2>&1; grep "tclust timing\|real\|time of\|ERROR" log$it.txt; echo ---; done
