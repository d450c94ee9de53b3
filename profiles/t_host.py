import os, sys, numpy as np, time
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
os.environ['PLG_TIMING']='1'
from util import random_desc
from pylogeny_b200 import fitch, neighbourhood as nbh, likelihood
n,S=256,100000
rng=np.random.default_rng(1002)
states=(rng.integers(0,4,(S,n))+48).astype(np.uint8)
a=fitch.FitchAlignment.from_dense(states,np.ones(S,dtype=np.int64))
d=random_desc(np.random.default_rng(5),n)
for i in range(3):
    t=time.perf_counter(); m,c=nbh.score_parsimony(a,d); print('total',(time.perf_counter()-t)*1e3, file=sys.stderr)
print('--- lnL 64 x 10k', file=sys.stderr)
n,S=64,10000
rng=np.random.default_rng(1001)
codes=(1<<rng.integers(0,4,(n,S))).astype(np.uint8)
la=likelihood.LikAlignment(codes,np.ones(S))
m=likelihood.Model(np.array([1.2,3.4,0.8,1.1,4.1,1.0]),np.array([0.28,0.22,0.24,0.26]),0.7)
d=random_desc(np.random.default_rng(6),n); br=np.random.default_rng(7).exponential(0.1,d.shape)
for i in range(3):
    t=time.perf_counter(); mv,l=nbh.score_likelihood(la,m,d,br); print('total',(time.perf_counter()-t)*1e3, file=sys.stderr)
