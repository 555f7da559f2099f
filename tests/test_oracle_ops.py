"""CPU: the per-layer reference (oracle/port_ops.py) reproduces the layer taps of the pinned whole-path port
(oracle/port.py, itself checked against the unmodified reference's goldens in test_oracle_golden.py)."""
import torch

from oracle import port, port_ops


def test_layer_reference_matches_port_taps():
    pcfg = port.PathConfig(proc_side=64)
    spec = port.effnet_spec('efficientnetv2-tiny')
    sd = port.make_effnet_state_dict(spec, pcfg, 8, seed=0)
    crops, _ = port.synthetic_inputs(2, 64, seed=0)
    tap = {}
    with torch.inference_mode():
        port.effnet_features(sd, spec, crops, tap=tap)
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
    P = 'backbone.1'
    cases = [
        (f'{P}.0', crops, None, tap[f'{P}.0']),                                        # stem (NCHW crops in)
        (f'{P}.1.0.block.0', nhwc(tap[f'{P}.0']), nhwc(tap[f'{P}.0']), tap[f'{P}.1.0']),  # expand-1 fused block + residual
        (f'{P}.2.0.block.0', nhwc(tap[f'{P}.1.0']), None, tap[f'{P}.2.0.block.0']),    # 3x3 stride 2
        (f'{P}.2.1.block.1', nhwc(tap[f'{P}.2.1.block.0']), nhwc(tap[f'{P}.2.0']), tap[f'{P}.2.1']),  # project + residual
        (f'{P}.4.0.block.0', nhwc(tap[f'{P}.3.0']), None, tap[f'{P}.4.0.block.0']),    # MBConv expand
        (f'{P}.4.0.block.1', nhwc(tap[f'{P}.4.0.block.0']), None, tap[f'{P}.4.0.block.1']),  # depthwise stride 2
        (f'{P}.7', nhwc(tap[f'{P}.6.1']), None, tap[f'{P}.7']),                        # last conv
    ]
    for name, x, res, want in cases:
        got = port_ops.conv_layer_reference(sd, spec, name, x, res, dtype=torch.float64).permute(0, 3, 1, 2)
        assert port.relative_error(got, want) < 2e-6, name
    # squeeze-excitation projection: scale * x then 1x1 conv (efficientnet.py:110-173)
    key = f'{P}.4.0.block'
    dw = tap[f'{key}.1']
    s = dw.mean(dim=(2, 3), keepdim=True)
    s = torch.nn.functional.silu(torch.nn.functional.conv2d(s, sd[f'{key}.2.fc1.weight'], sd[f'{key}.2.fc1.bias']))
    s = torch.sigmoid(torch.nn.functional.conv2d(s, sd[f'{key}.2.fc2.weight'], sd[f'{key}.2.fc2.bias']))
    got = port_ops.conv_layer_reference(sd, spec, f'{key}.3', nhwc(dw), None, scale=s[:, :, 0, 0], dtype=torch.float64)
    assert port.relative_error(got.permute(0, 3, 1, 2), tap[f'{key}.3']) < 2e-6


def test_bf16_rounding_points():
    """'bf16' precision rounds the folded weight once and the (scaled) input once; everything else stays wide."""
    pcfg = port.PathConfig(proc_side=64)
    spec = port.effnet_spec('efficientnetv2-tiny')
    sd = port.make_effnet_state_dict(spec, pcfg, 8, seed=0)
    x = torch.randn(2, 16, 16, 16).bfloat16().float()
    a = port_ops.conv_layer_reference(sd, spec, 'backbone.1.3.0.block.0', x, precision='bf16')
    b = port_ops.conv_layer_reference(sd, spec, 'backbone.1.3.0.block.0', x, precision='exact')
    err = port.relative_error(a, b)
    assert 1e-5 < err < 2e-2  # differs by the weight rounding only
