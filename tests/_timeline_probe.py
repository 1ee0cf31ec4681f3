import numpy as np, go_ctr_b200 as g
from tests.util import make_batch
B=65536
cfg=g.engine.default_config(g.MODEL_DIN_COS,uP=52,S=50,D=64,cF=53,batch=B,pred_batch=B)
e=g.Engine(cfg)
e.table_fill(0,1000,52); e.table_fill(1,5000,53); e.table_fill(2,5000,64,dist=1,scale=.125)
rng=np.random.default_rng(0)
ur,ir,hist,y=make_batch(rng,1000,5000,B,50)
for i in range(2): e.train_step_idx(ur,ir,hist,y)
