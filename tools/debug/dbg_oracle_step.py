import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import c_oracle, np_oracle as O
from pypic_b200.rng import LegacyDraws
from pypic_b200.sheath import SheathSim
def run(N, Ng, sort_every, rng, track=None, E0amp=0.0):
    dx, dt = 1e-5, 1e-12
    L = dx*(Ng-1); kT = O.kb*116000.; h=N//2
    rs=np.random.RandomState(11)
    x0=rs.uniform(0,L,N)
    sig=np.concatenate([np.full(h,np.sqrt(kT/O.me)),np.full(N-h,np.sqrt(kT/O.mp))])
    u0=rs.normal(0,1,N)*sig
    E0=rs.normal(0,E0amp,Ng) if E0amp else np.zeros(Ng); p2c=L*1e19/N
    act=np.ones(N)
    x1,u1,E1,j1,k,r=c_oracle.dd_picard_step(x0,u0,[-O.e,O.e],[O.me,O.mp],h,act,E0,p2c,Ng,dx,dt,L,1e-5,20,16)
    sim=SheathSim(N,Ng,dx,dt,p2c,kBT=(kT,kT),carry_vw=False,rng=rng,sort_every=sort_every,draws=LegacyDraws(np.random.RandomState(77)),track_order=track)
    sim.upload(x0,u0,E0=E0)
    kg,rg=sim.step()
    out=sim.download()
    dE=out["E0"]-E1; dj=out["j0"]-j1
    print("N=%g Ng=%d sort=%d rng=%s track=%s: k %d/%d dead %d flags_eq %s | dE max %.3e at %d, mean %.3e | dj max %.3e at %d, sum %.3e | j edges gpu %s cpu %s" % (
        N,Ng,sort_every,rng,sim.track,kg,k,(act!=1).sum(), np.array_equal(out["active"] if sim.oid is None or True else 0,act),
        np.abs(dE).max(), np.abs(dE).argmax(), dE.mean(), np.abs(dj).max(), np.abs(dj).argmax(), dj.sum(),
        out["j0"][[0,1,-2,-1]], j1[[0,1,-2,-1]]))
    sys.stdout.flush()
for args in [(1000000,4097,0,"host"),(1000000,4097,2,"host"),(1000000,257,2,"host"),(10000000,4097,0,"host"),(10000000,4097,2,"host"),(1000000,4097,2,"philox",False)]:
    run(*args)
