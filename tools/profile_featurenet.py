"""One FeatureNet forward (5 DTU views, native engine) between cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmvsnet_b200 import MVSNet, synthetic as syn
net = MVSNet([48, 32, 8], [4, 2, 1]).cuda().eval()
imgs = syn.make_images(1184, 1600, 5, 1, natural=True).cuda()
with torch.no_grad():
    for _ in range(3):
        net.feature(imgs[0])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    net.feature(imgs[0])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
