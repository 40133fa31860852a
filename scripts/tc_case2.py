import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, density_planner_b200 as dp
N, T, stride, margin, impl = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), (sys.argv[5] if sys.argv[5] != "none" else None)
args = dp.default_args(); car = dp.Car(args)
torch.manual_seed(3)
up, uref, xref = car.get_valid_ref(args)
uref, xref = uref[:, :, :T], xref[:, :, :T + 1]
xe0 = car.sample_xe0(N).cuda()
print("launch", flush=True)
cut = bool(int(os.environ.get("CUT", "1")))
out = car.compute_density(xe0, xref, uref, args.dt_sim, log_density=True, impl=impl, return_margin=bool(margin), stride=stride, cutting=cut)
torch.cuda.synchronize()
print("OK N=%d T=%d stride=%d margin=%d impl=%s rho[-1]=%.4f" % (N, T, stride, margin, impl, float(out[1][0, 0, -1])), flush=True)
