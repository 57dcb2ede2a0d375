import time; t0=time.perf_counter()
import sys, os; sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np; t1=time.perf_counter()
from impop_b200 import _native; lib=_native.lib(); t2=time.perf_counter()
from impop_b200.engine import Context
ctx=Context(0, lite=True); t3=time.perf_counter()
a=ctx.upload(np.random.rand(466,466)); t4=time.perf_counter()
st,ct,ws=ctx.reduce_identity(a, None, None); ctx.check(); t5=time.perf_counter()
print("numpy import %.3f lib load %.3f context %.3f upload %.3f reduce %.3f total %.3f" % (t1-t0,t2-t1,t3-t2,t4-t3,t5-t4,t5-t0))
