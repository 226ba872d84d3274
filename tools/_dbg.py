import sys
sys.path.insert(0, "/root/repo")
from ldbc_graphalytics_platforms_graphblas_b200 import capi
capi.init(0)
g = capi.Graph.rmat(20, False, weighted=False, want_mapping=False)
g.cdlp(10, out=False)
capi.profile(True)
g.cdlp(10, out=False)
capi.profile(False)
for k, v in capi.profile_report().items(): print(k, v)
