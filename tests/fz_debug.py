import sys, os, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, util, ddc_svd_b200 as D
for (m, n, env) in [(300,200,{"SVD_GPU_FUSED_MIN_ROWS":"8","SVD_GPU_FUSED_MIN_COLS":"3"}), (1500,1400,{}), (4200,300,{}), (8704,256,{}), (16500,130,{}), (33000,100,{})]:
    A = util.rand_matrix(m, n, 1.0, 2.0, 4)
    Ao, ao, bo = util.oracle_bidiag(A)
    os.environ.update(env)
    t=time.time(); Ag, ag, bg = D.bidiag_par(A); dt=time.time()-t
    for k_ in env: del os.environ[k_]
    print((m,n), "A err", np.abs(Ag-Ao).max(), "alpha", np.abs(ag-ao).max(), "beta", np.abs(bg-bo).max(), "nan", int(np.isnan(Ag).sum()), "t", round(dt,3), flush=True)
