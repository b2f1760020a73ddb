"""The shipped model configuration (configs/infer/DeepPointMap_B_Main_SemanticKITTI.yaml:32-60
and :66 `coor_scale`) as a plain namespace -- what `Encoder(args)` / `Decoder(args)` read
(network/encoder/encoder.py:14-22, network/decoder/decoder.py:15-21,217)."""
from types import SimpleNamespace

DPM_B = dict(
    encoder=dict(
        npoint=[4096, 1024, 256, 64, 16],
        radius_list=[[0.05, 0.1], [0.1, 0.2], [0.2, 0.4, 0.4], [0.4, 0.8], [0.8, 1.6]],
        nsample_list=[[32, 32], [32, 32], [32, 32, 32], [32, 32], [16, 16]],
        in_channel=3, out_channel=128, width=16, expansion=4, upsample_layers=2,
        sample=[dict(type="fps-t3d")] * 5, norm="LN", bias=True,
    ),
    decoder=dict(in_channel=128, model_channel=256, attention_layers=3),
    loss=dict(tau=0.1, eps_offset=2.0),
    coor_scale=60.0,
)


def _ns(d):
    if isinstance(d, dict):
        return SimpleNamespace(**{k: _ns(v) for k, v in d.items()})
    if isinstance(d, list):
        return [_ns(v) for v in d]
    return d


def dpm_b_config() -> SimpleNamespace:
    return _ns(DPM_B)
