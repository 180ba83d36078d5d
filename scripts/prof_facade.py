import cProfile, pstats, sys, torch
sys.path.insert(0, '/root/repo')
import bench
pr = cProfile.Profile()
bench.e2e_facade("sri_resnet_a", 512, 1, torch.device('cuda', 0))
pr.enable()
r = bench.e2e_facade("sri_resnet_a", 512, 3, torch.device('cuda', 0))
pr.disable()
print(r)
pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
