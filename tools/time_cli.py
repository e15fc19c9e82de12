import subprocess, time, os, sys
os.chdir("oracle/_ref")
def run(cmd, env=None):
    t=time.time(); r=subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, **(env or {}))); dt=time.time()-t
    return dt, r
for name, cmd, env in [
    ("RNAcode_cuda gpu-evolve", ["./RNAcode_cuda","--gtf","--best-only","-n","1000","examples/genomic.maf"], {}),
    ("RNAcode_cuda host-evolve", ["./RNAcode_cuda","--gtf","--best-only","-n","1000","examples/genomic.maf"], {"RNACODE_CUDA_EVOLVE":"host"}),
    ("RNAcode_b200 pipeline", ["./RNAcode_b200","--gtf","--best-only","-n","1000","examples/genomic.maf"], {"RNACODE_CUDA_VERBOSE":"1"}),
    ("RNAcode_b200 pipeline (2nd)", ["./RNAcode_b200","--gtf","--best-only","-n","1000","examples/genomic.maf"], {"RNACODE_CUDA_VERBOSE":"1"}),
]:
    dt,r=run(cmd,env)
    print("%-30s %.2f s rc=%d lines=%d"%(name,dt,r.returncode,len(r.stdout.splitlines())))
    for l in r.stderr.splitlines():
        if "RNAcode_b200" in l: print("   ",l)

# synthetic MAF through the batched pipeline (config 3 shape, reduced block count)
sys.path.insert(0, "../..")
from rnacode_b200 import synth
for nblk, N, cols, n in [(2000, 10, 120, 100), (200, 10, 1000, 1000)]:
    blocks = [synth.synth_block(1, i, N, cols) for i in range(nblk)]
    path = "/tmp/synth_%d_%d_%d.maf" % (nblk, N, cols)
    synth.to_maf(blocks, path)
    for rep in range(2):
        dt, r = run(["./RNAcode_b200", "--tabular", "-n", str(n), path], {"RNACODE_CUDA_VERBOSE": "1"})
        print("pipeline %d blocks x %d x %d, -n %d: %.2f s (%.1f blocks/s) rc=%d out-lines=%d" % (nblk, N, cols, n, dt, nblk / dt, r.returncode, len(r.stdout.splitlines())))
        for l in r.stderr.splitlines():
            if "RNAcode_b200" in l: print("   ", l)
    if nblk == 2000:
        sub = "/tmp/synth_sub.maf"
        synth.to_maf(blocks[:32], sub)
        dt, r = run(["./RNAcode_ref", "--tabular", "-n", str(n), sub])
        print("reference, 32 of those blocks, one core: %.2f s (%.2f blocks/s)" % (dt, 32 / dt))
