import sys, os; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
os.environ['WBX_FIR']='tc'
import numpy as np, oracle_api as o, scenarios as sc, whitebox_b200 as wb
taps=int(sys.argv[1]) if len(sys.argv)>1 else 777
ref=sc.reverb(lambda C,B,r,bpm: o.Session('port',C,B,r,bpm), wb.effect_params, taps)
res=sc.reverb(lambda C,B,r,bpm: wb.Engine(C,B,r,bpm,device=0,sum_mode=wb.SUM_EXACT), wb.effect_params, taps)
peak=np.abs(ref['out']).max(axis=(1,2),keepdims=True)
err=np.abs(res['out'].astype(np.float64)-ref['out'])
print('max err / peak', float((err/peak).max()))
