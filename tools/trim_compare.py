import sys,time,numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import f16_mpc_oop_py_b200 as f16
f16.init()
hh,vv=np.meshgrid(np.linspace(5000,40000,64),np.linspace(300,900,64),indexing="ij")
res={}
for name,mode,stg in (("strict",f16.MATH_STRICT,1),("fast_new",f16.MATH_FAST,1),("fast_old",f16.MATH_FAST,0)):
    f16.lib.f16_set_math_mode(mode); f16.lib.f16_set_table_staging(stg)
    x,opt=f16.trim(hh.ravel(),vv.ravel(),fi=1,xcg=0.25)
    res[name]=(x,opt)
s=res["strict"]
for k in ("fast_new","fast_old"):
    x,opt=res[k]
    both=(s[1]["success"]>0)&(opt["success"]>0)
    cs,cf=s[1]["fun"][both],opt["fun"][both]
    print(k,"both converged",both.mean(),"cost ratio pct [5,50,95]",np.percentile(cf/np.maximum(cs,1e-300),[5,50,95]),
          "max |dalpha| deg",np.rad2deg(np.abs(x[7,both]-s[0][7,both]).max()),"max |dT|",np.abs(x[12,both]-s[0][12,both]).max(),
          "pct95 |dT|",np.percentile(np.abs(x[12,both]-s[0][12,both]),95), "pct95 |dalpha| deg", np.percentile(np.rad2deg(np.abs(x[7,both]-s[0][7,both])),95))
print("strict cost percentiles",np.percentile(s[1]["fun"],[5,25,50,75,95]))
