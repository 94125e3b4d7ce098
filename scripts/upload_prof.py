#!/usr/bin/env python
"""Host wall-clock breakdown of ba_gpu_upload over the cfg 2 sliding sequence through the compiled drop-in
(BA_UPLOAD_PROF=1 stamps on stderr; averaged here over the second pass)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import ba_b200
    for rep in range(2):
        seq = ba_b200.synthetic.make_config(2)
        sys.stderr.write("[PASS %d]\n" % rep)
        res = ba_b200.hostlib.sliding_sequence(seq, 20, 10, max_num_iterations=10, fixed_iterations=True)
    print(res["windows"], {k: round(v / res["windows"], 4) for k, v in res["ms"].items()})
    sys.exit(0)
env = dict(os.environ, BA_UPLOAD_PROF="1")
out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
print(out.stdout.strip())
lines = out.stderr.split("[PASS 1]")[-1].splitlines()
acc, n = {}, 0
for l in lines:
    if "[BA_UPLOAD_PROF]" not in l:
        continue
    n += 1
    for name, us in re.findall(r"([a-z0-9 +]+?) (\d+) us \|", l.split("]", 1)[1]):
        acc[name.strip()] = acc.get(name.strip(), 0.0) + float(us)
print("windows", n, {k: round(v / max(n, 1), 1) for k, v in acc.items()}, "us per upload")
