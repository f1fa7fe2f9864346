import os, sys, subprocess, json
sys.path.insert(0, '/root/repo')
import numpy as np
mode = sys.argv[1]
os.environ['SDA_PHASED'] = mode
import sdaopt_b200 as sb
f = sb.Rastrigin(50)
r = sb.sda_batch(f, f.bounds, 4096, maxiter=60)
np.savez('/tmp/ph_%s.npz' % mode, **{k: r[k] for k in ('x','fun','nit','ncall','ncall_ls','n_ls','status')})
print(mode, r.ncall.sum(), r.n_ls.sum(), r.fun.min(), (r.status!=0).sum())
