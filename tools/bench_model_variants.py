import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
from quadrotorilqr_b200 import BatchILQR, problems
from quadrotorilqr_b200.options import ILQROptions
m=problems.hover_model(); opts=problems.default_options(False)
B,N=int(sys.argv[1]) if len(sys.argv)>1 else 65536,40
desired=problems.hover_desired_trajectory(N,m["dt_s"],m["mass_kg"],m["g_mpss"])
x0=problems.hover_initial_states(B,seed=0)
seed=problems.constant_state_trajectory(x0,N,m["dt_s"],desired[0,14:18])
for flags in (0,4,1,3):
    s=BatchILQR(m["mass_kg"],m["inertia"],m["arm_length_m"],m["torque_to_thrust_ratio_m"],m["g_mpss"],m["Q"],m["R"],m["dt_s"],opts,model_flags=flags)
    init=s.forward_sim(seed,np.zeros((B,N,4)),np.zeros((B,N,48)))
    s.set_profiling(True)
    for rep in range(2):
        t=time.time(); r=s.solve(init,desired); dt=time.time()-t
    st=s.last_solve_stats()
    res=r["results"]
    print(flags,"host-API solve %.1f ms"%(dt*1e3),"conv",np.isin(res["status"],[1,2]).mean(),"iters",res["backward_passes"].mean(), {k:(round(v,2) if isinstance(v,float) else v) for k,v in st.items()} if isinstance(st,dict) else st, flush=True)
