import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from oracle import pinocchio_oracle as po
from pinocchio_b200.cosmology import Cosmology, SmoothingLadder
from pinocchio_b200.engine import Pinocchio, RunConfig
np.set_printoptions(precision=17, linewidth=200)
R=[20.635922, 13.996056, 9.026099, 5.465945, 3.058354, 1.548258, 0.689079, 0.258729, 0.0]
cosmo=Cosmology(pk_norm_override=2.03146e7)
N=64
p=Pinocchio(RunConfig(GridSize=N,BoxSize_htrue=N/0.7), cosmo, smoothing=SmoothingLadder(np.array(R),np.zeros(9)))
p.GenIC_large(); kd=p.read_kdensity()
tot=0
for ism,r in enumerate(R):
    h=p.compute_second_derivatives(r)
    F=p.inverse_collapse_time(h)
    ref=po.inverse_collapse_time([h[i] for i in range(6)], cosmo.InverseGrowingMode)
    d=np.abs(F-ref)
    bad=d>1e-9*np.maximum(1,np.abs(ref))
    print('radius',ism,r,'nbad',bad.sum(),'max',d.max(), 'nan',np.isnan(F).sum(), np.isnan(ref).sum())
    if bad.any():
        idx=np.argwhere(bad)[:5]
        for i in idx:
            i=tuple(i)
            hh=np.array([h[c][i] for c in range(6)])
            x1,x2,x3,bd=po.eigenvalues([np.array([v]) for v in hh])
            bc=po.ell_classic(x1,x2,x3)
            print('  cell',i,'h',hh,'F gpu',F[i],'ref',ref[i],'eig',x1,x2,x3,'bc',bc)
# full path
p.compute_fmax(displacements=False)
ref=po.compute_fmax(kd,R,1/0.7,cosmo.InverseGrowingMode,lpt_order=0,keep=True)
Fm=p.field('Fmax'); Rm=p.field('Rmax')
d=np.abs(Fm.astype(np.float64)-ref['Fmax'])
bad=d>1e-6*np.maximum(1,np.abs(ref['Fmax']))
print('fmax nbad',bad.sum(),'max',d.max(),'rmax mismatch',(Rm!=ref['Rmax']).sum())
for i in np.argwhere(bad)[:10]:
    i=tuple(i); print(' ',i,Fm[i],ref['Fmax'][i],Rm[i],ref['Rmax'][i],[f[i] for f in ref['F']])
# random synthetic
rng=np.random.default_rng(3); n=200000
h=rng.standard_normal((6,n))*np.array([1.5,1.5,1.5,.7,.7,.7])[:,None]
F=p.inverse_collapse_time(h); ref=po.inverse_collapse_time([h[i] for i in range(6)],cosmo.InverseGrowingMode)
d=np.abs(F-ref); bad=~(d<=1e-10*np.maximum(1,np.abs(ref)))
print('synthetic nbad',bad.sum(), 'nan',np.isnan(F).sum(),np.isnan(ref).sum())
for i in np.argwhere(bad)[:10,0]:
    x1,x2,x3,bd=po.eigenvalues([h[c][i:i+1] for c in range(6)])
    print('  ',i,h[:,i],F[i],ref[i],x1,x2,x3,po.ell_classic(x1,x2,x3))
