"""CPU: the bf16-storage restatement (oracle/port_bf16.py) is the pinned fp32 port plus roundings - with the roundings
switched off it reproduces the port; with them on it quantifies what bf16 STORAGE alone does to the reference arithmetic on
the conditioned random weights the parity tests use (no kernel of this repo involved)."""
import torch

from oracle import port, port_bf16


def _setup(name, side, j, batch):
    pcfg = port.PathConfig(proc_side=side)
    spec = port.effnet_spec(name)
    sd = port.make_effnet_state_dict(spec, pcfg, j, seed=0, calib_batch=2)
    crops, k = port.synthetic_inputs(batch, side, seed=0)
    return pcfg, spec, sd, crops, k


def test_restatement_without_rounding_equals_the_port(monkeypatch):
    """Folding batch norm into the weights (what the engine does at load) and adding the residual before the final
    rounding point is the same function as the port's conv -> BN -> SiLU chain."""
    pcfg, spec, sd, crops, k = _setup('efficientnetv2-tiny', 64, 8, 3)
    monkeypatch.setattr(port_bf16, '_q', lambda x: x)
    with torch.inference_mode():
        s_ref, s_new = {}, {}
        ref = port.metrabs_forward(sd, spec, pcfg, 8, crops, k, stages=s_ref)
        out = port_bf16.metrabs_forward_bf16(sd, spec, pcfg, 8, crops, k, stages=s_new)
    assert port.relative_error(s_new['features'], s_ref['features']) < 2e-5
    assert port.relative_error(out, ref) < 5e-3   # fp32 reassociation of the BN fold, amplified by the decode of this tiny net


def test_bf16_storage_alone_exceeds_the_1e3_bar_on_untrained_weights():
    """the 1e-3 bar is held in fp32 mode; ANY bf16-storage evaluation of this untrained network -
    here the reference arithmetic on the CPU with roundings inserted - is off by percent on the features and by tens of
    percent on the joints, which is the scale the device's bf16 mode is compared with (tests/test_gpu_tc.py)."""
    pcfg, spec, sd, crops, k = _setup('efficientnetv2-s', 256, 24, 2)
    with torch.inference_mode():
        s_ref, s_b = {}, {}
        ref = port.metrabs_forward(sd, spec, pcfg, 24, crops, k, stages=s_ref)
        out = port_bf16.metrabs_forward_bf16(sd, spec, pcfg, 24, crops, k, stages=s_b)
    e_feat = port.relative_error(s_b['features'], s_ref['features'])
    e_out = port.relative_error(out, ref)
    print(f'bf16 storage on the CPU restatement vs fp32: features {e_feat:.3e}, joints {e_out:.3e}')
    assert torch.isfinite(out).all()
    assert 1e-3 < e_feat < 0.5      # percent-level on the features ...
    assert e_out > 1e-2             # ... amplified well past the 1e-3 bar on the joints
