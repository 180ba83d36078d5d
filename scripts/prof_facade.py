"""cProfile of NetworkAbstractor.forward through the reference-API facade (bench.e2e_facade)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
pr = cProfile.Profile()
bench.e2e_facade("sri_resnet_a", 512, 1, torch.device('cuda', 0))
pr.enable()
r = bench.e2e_facade("sri_resnet_a", 512, 3, torch.device('cuda', 0))
pr.disable()
print(r)
pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
