import sys, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo')
from openstereo_b200 import ops
b,cin,cout,d,h,w = [int(a) for a in sys.argv[1:7]]
g = torch.Generator().manual_seed(0)
x, wt = torch.randn(b,cin,d,h,w, generator=g), torch.randn(cin,cout,3,3,3, generator=g)*0.2
sc, sh = torch.rand(cout, generator=g)+0.5, torch.randn(cout, generator=g)*0.1
want = F.conv_transpose3d(x.double(), wt.double(), stride=2, padding=1, output_padding=1).float()
res = torch.randn(*want.shape, generator=g)
want2 = F.relu(want*sc.view(1,-1,1,1,1)+sh.view(1,-1,1,1,1)+res)
xc = ops.to_ndhwc(x.cuda()); wp = ops.pack_tc_deconv_weight(wt.cuda())
def rel(a,b): return ((a.cpu()-b).abs().max()/b.abs().max()).item()
got = ops.deconv3d_k3_tc(xc, wp); torch.cuda.synchronize(); print('plain', rel(got,want), flush=True)
got = ops.deconv3d_k3_tc(xc, wp, sc.cuda(), sh.cuda(), res.cuda(), ops.ACT_RELU); torch.cuda.synchronize(); print('res ncdhw', rel(got,want2), flush=True)
got = ops.deconv3d_k3_tc(xc, wp, sc.cuda(), sh.cuda(), res.permute(0,2,3,4,1).contiguous().cuda(), ops.ACT_RELU, out_ndhwc=True, res_ndhwc=True); torch.cuda.synchronize(); print('ndhwc', rel(got.permute(0,4,1,2,3),want2), flush=True)
