import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[os.path.join(ROOT,'spmv-fpga_b200'), os.path.join(ROOT,'tests')]
import numpy as np, spmvb, oracle_api as oa
O=oa.OracleLib()
nx=int(sys.argv[1]) if len(sys.argv)>1 else 2048
A=spmvb.Csr.laplacian2d(nx,nx)
lay=spmvb.Layout.from_csr(A)
x=np.random.default_rng(5).random(A.cols)
gold=O.spmv_gold(A.rows,A.row_ptr,A.col_ind,A.values,x,True)
for v in (1,2,4,5):
    eng=spmvb.Engine(lay,0,v)
    for rep in range(3):
        y=np.zeros(A.rows); eng.spmv_host(x,y,accumulate=False)
        bad=np.nonzero(np.abs(y-gold)>1e-9)[0]
        print('variant',v,'rep',rep,'bad rows',len(bad), bad[:12], (y-gold)[bad[:6]])
        if len(bad):
            b=bad[0]; print('   row',b,'block',b//32768,'chunk approx', b*5//256, 'gold',gold[b],'got',y[b])
    eng.free()
