"""Drop-in replacements of the reference's three CPython extension modules (same names, same positional
signatures, caller-allocated outputs).  `pointrcnn_b200.dropin.activate()` registers them in sys.modules
as `pointnet2_cuda`, `iou3d_cuda`, `roipool3d_cuda`."""
