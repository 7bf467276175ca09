import time, io, contextlib, sys, os
sys.path.insert(0, os.getcwd())
from kmerpapa_b200 import synthetic, cli
kmers, pos, neg = synthetic.negbin_counts("NNNNANNNN", 9004)
with open("/tmp/pos9.txt","w") as f:
    for k,p in zip(kmers,pos): f.write(f"{k} {p}\n")
with open("/tmp/bg9.txt","w") as f:
    for k,p,n in zip(kmers,pos,neg): f.write(f"{k} {p+n}\n")
for rep in range(2):
    t=time.perf_counter()
    err=io.StringIO()
    with contextlib.redirect_stderr(err):
        cli.main(["-p","/tmp/pos9.txt","-b","/tmp/bg9.txt","-c","3","5","6","-a","0.5","1","10","--nfolds","5","--seed","1","-o","/tmp/out9.txt","--CVfile","/tmp/cv9.txt"])
    print("cfg4 CLI end to end", round(time.perf_counter()-t,2), "s")
print(err.getvalue()[-600:])
print(open("/tmp/cv9.txt").read())
import cProfile, pstats
pr=cProfile.Profile(); pr.enable()
with contextlib.redirect_stderr(io.StringIO()):
    cli.main(["-p","/tmp/pos9.txt","-b","/tmp/bg9.txt","-c","3","5","6","-a","0.5","1","10","--nfolds","5","--seed","1","-o","/tmp/out9.txt"])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
