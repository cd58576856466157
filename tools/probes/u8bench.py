import sys, json
sys.path.insert(0, '/root/repo/tools'); sys.path.insert(0, '/root/repo')
import microbench as mb
for wm in (True, False):
    r = mb.bench_gather_u8(512, 256, 3, 15, 8, with_mean=wm)
    print(wm, round(r['ms_epoch'], 3), round(r['gbs']))
