import subprocess, time, os, sys
os.chdir("oracle/_ref"); sys.path.insert(0, "../..")
from rnacode_b200 import synth
blocks = [synth.synth_block(1, i, 10, 120) for i in range(2000)]
synth.to_maf(blocks, "/tmp/s.maf")
for rep in range(2):
    r = subprocess.run(["./RNAcode_b200", "--tabular", "-n", "100", "/tmp/s.maf"], capture_output=True, text=True, env=dict(os.environ, RNACODE_CUDA_VERBOSE="1"))
    print("\n".join(l for l in r.stderr.splitlines() if "RNAcode_b200" in l))
