import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b3d = importlib.import_module("3d-brain-tumor-segmentation_b200")
ops = b3d.ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
def run(shape, cin, cout, ones=False):
    x = torch.ones(1, *shape, cin, device=dev) if ones else torch.randn(1, *shape, cin, device=dev)
    dy = torch.ones(1, *shape, cout, device=dev) if ones else torch.randn(1, *shape, cout, device=dev)
    out = []
    for tc in (0, 1):
        dw = torch.full((3, 3, 3, cin, cout), 7.0, device=dev)
        try:
            xb = torch.empty(x.shape, device=dev, dtype=torch.bfloat16) if tc else None
            yb = torch.empty(dy.shape, device=dev, dtype=torch.bfloat16) if tc else None
            ops._call("b3d_conv3d_wgrad", x, dy, dw, None, 1, 0, xb, yb)
            torch.cuda.synchronize()
        except Exception as e:
            print("EXC", tc, e)
        out.append(dw)
    a, b = out
    err = float((a - b).norm() / a.norm())
    print(shape, cin, cout, "rel", err, "tc absmax", float(b.abs().max()), "ref absmax", float(a.abs().max()),
          "nan", bool(torch.isnan(b).any()))
    if err > 1e-2:
        print(" ref[1,1,1,:4,:4]\n", a[1, 1, 1, :4, :4].cpu().numpy())
        print(" tc [1,1,1,:4,:4]\n", b[1, 1, 1, :4, :4].cpu().numpy())
        print(" tc [0,0,0,:4,:4]\n", b[0, 0, 0, :4, :4].cpu().numpy())
run((2, 4, 8), 16, 16, ones=True)
run((2, 4, 8), 16, 16)
run((4, 8, 16), 16, 16)
run((4, 8, 16), 32, 32)
run((4, 8, 16), 64, 64)
run((4, 8, 16), 128, 128)
run((5, 7, 9), 192, 64)
