import sys
sys.path.insert(0, ".")
from umgen_b200 import capi
import torch
torch.cuda.init()
torch.zeros(1, device="cuda")
lib = capi.lib()
print("capacity", lib.umgen_decode_cluster_capacity(), lib.umgen_last_error())
