import importlib, os, sys, torch, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b3d = importlib.import_module("3d-brain-tumor-segmentation_b200")
ops = b3d.ops
dev = torch.device("cuda:0")
dbg = torch.full((128,), -1.0, device=dev)
b3d._lib.lib.b3d_debug_wgrad_buffer.argtypes = [ctypes.c_void_p]
b3d._lib.lib.b3d_debug_wgrad_buffer(dbg.data_ptr())
shape, cin, cout = (2, 4, 8), 16, 16
x = torch.ones(1, *shape, cin, device=dev) * 2
dy = torch.ones(1, *shape, cout, device=dev) * 3
dw = torch.full((3, 3, 3, cin, cout), 7.0, device=dev)
ops._call("b3d_conv3d_wgrad", x, dy, dw, None, 1, 0, 1)
torch.cuda.synchronize()
d = dbg.cpu().numpy()
print("xs", d[:8], "ys", d[8:16], "params", d[16:22])
print("tmem tap13 lanes0-3:\n", d[32:96].reshape(4, 16))
print("dw absmax", float(dw.abs().max()))
