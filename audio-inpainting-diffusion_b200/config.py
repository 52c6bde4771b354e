"""Configuration of the denoiser path, read from the reference's `args` object.

Mirrors what `Unet_CQT_oct_with_attention.__init__` reads (unet.py:595-655): `args.network.*` and
`args.exp.{sample_rate,audio_len}`.  `args` may be an OmegaConf/EasyDict-style attribute object or nested dicts.
"""
from dataclasses import dataclass, field, asdict
from typing import List

from . import _lib

_MISSING = object()


def cfg_get(obj, path, default=_MISSING):
    cur = obj
    for part in path.split("."):
        try:
            cur = cur[part] if isinstance(cur, dict) else getattr(cur, part)
        except (KeyError, AttributeError):
            if default is _MISSING:
                raise KeyError(f"config entry '{path}' is missing") from None
            return default
    return cur


class AttrDict(dict):
    """Minimal attribute dict (stand-in for the reference's OmegaConf objects; utils/dnnlib/util.py:39-52)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(d):
        if isinstance(d, dict):
            return AttrDict({k: AttrDict.wrap(v) for k, v in d.items()})
        return d


@dataclass
class NetConfig:
    num_octs: int = 7
    bins_per_oct: int = 64
    sample_rate: float = 22050.0
    audio_len: int = 262144
    window: str = "kaiser"
    beta: float = 1.0
    emb_dim: int = 256
    Ns: List[int] = field(default_factory=lambda: [64, 96, 96, 128, 128, 256, 256])
    num_dils: List[int] = field(default_factory=lambda: [2, 3, 4, 5, 6, 7, 7])
    attention_layers: List[int] = field(default_factory=lambda: [0, 0, 0, 0, 1, 1, 1, 1])
    num_heads: int = 8
    num_bottleneck_layers: int = 1
    conv_mode: int = 0  # 0 = exact fp32 CUDA cores; 1 = tcgen05 with split-fp16 operands (fp32-grade, 3 MMAs per tap); 2 = tcgen05, single fp16 operands

    @staticmethod
    def from_args(args, conv_mode=None):
        """conf/network/paper_1912_unet_cqt_oct_attention_adaLN_2.yaml + conf/exp/*.yaml layout."""
        g = lambda p, d=_MISSING: cfg_get(args, p, d)
        if g("network.use_fencoding", False):
            raise NotImplementedError("use_fencoding=True is not on the hot path (yaml:6)")
        if not g("network.use_norm", True):
            raise NotImplementedError("use_norm=False is not supported")
        if g("network.bottleneck_type", "res_dil_convs") != "res_dil_convs":
            raise NotImplementedError("bottleneck type not implemented")  # unet.py:694
        if g("network.attention_dict.use_rel_pos", False) or g("network.attention_dict.bias_qkv", False):
            raise NotImplementedError("use_rel_pos / bias_qkv are off in the reference configs and not supported")
        window = g("network.cqt.window", "kaiser")
        cfg = NetConfig(
            num_octs=int(g("network.cqt.num_octs")), bins_per_oct=int(g("network.cqt.bins_per_oct")),
            sample_rate=float(g("exp.sample_rate")), audio_len=int(g("exp.audio_len")),
            window=str(window), beta=float(g("network.cqt.beta", 1.0)), emb_dim=int(g("network.emb_dim")),
            Ns=[int(v) for v in g("network.Ns")], num_dils=[int(v) for v in g("network.num_dils")],
            attention_layers=[int(v) for v in g("network.attention_layers")],
            num_heads=int(g("network.attention_dict.num_heads", 8)),
            num_bottleneck_layers=int(g("network.num_bottleneck_layers", 1)),
            conv_mode=int(g("network.conv_mode", 0)) if conv_mode is None else int(conv_mode))
        cfg.validate()
        return cfg

    def validate(self):
        no = self.num_octs
        if not (1 <= no <= _lib.MAX_OCTS):
            raise ValueError("num_octs out of range")
        if len(self.Ns) < no or len(self.num_dils) < no or len(self.attention_layers) < no + 1:
            raise ValueError("Ns / num_dils need num_octs entries and attention_layers num_octs+1")
        if self.window not in ("kaiser", "hann"):
            raise ValueError(f"unsupported CQT window {self.window!r}")

    def to_c(self):
        self.validate()
        c = _lib.AidConfig()
        c.num_octs, c.bins_per_oct, c.audio_len = self.num_octs, self.bins_per_oct, self.audio_len
        c.window_kind = 1 if self.window == "kaiser" else 0
        c.sample_rate, c.beta = self.sample_rate, self.beta
        c.emb_dim, c.num_heads = self.emb_dim, self.num_heads
        for i in range(self.num_octs):
            c.Ns[i], c.num_dils[i] = self.Ns[i], self.num_dils[i]
        for i in range(self.num_octs + 1):
            c.attention_layers[i] = self.attention_layers[i]
        c.num_bottleneck_layers, c.conv_mode = self.num_bottleneck_layers, self.conv_mode
        return c

    def to_args(self):
        """The `args` object the reference constructor would take for this configuration."""
        return AttrDict.wrap({
            "exp": {"sample_rate": int(self.sample_rate), "audio_len": self.audio_len},
            "network": {
                "cqt": {"num_octs": self.num_octs, "bins_per_oct": self.bins_per_oct, "window": self.window, "beta": self.beta},
                "emb_dim": self.emb_dim, "use_norm": True, "use_fencoding": False, "Ns": list(self.Ns),
                "Ss": [2] * self.num_octs, "num_dils": list(self.num_dils), "attention_layers": list(self.attention_layers),
                "bottleneck_type": "res_dil_convs", "num_bottleneck_layers": self.num_bottleneck_layers,
                "conv_mode": self.conv_mode,
                "attention_dict": {"num_heads": self.num_heads, "attn_dropout": 0.0, "bias_qkv": False, "N": 0,
                                   "rel_pos_num_buckets": 32, "rel_pos_max_distance": 64, "use_rel_pos": False, "Nproj": 8},
            },
        })

    def as_dict(self):
        return asdict(self)


def paper_22k(audio_len=262144, conv_mode=0):
    """conf/network/paper_1912_unet_cqt_oct_attention_adaLN_2.yaml at fs=22050 (the BASELINE network)."""
    return NetConfig(audio_len=audio_len, conv_mode=conv_mode)


def paper_44k(audio_len=184184, conv_mode=0):
    """conf/network/paper_1912_unet_cqt_oct_attention_44k_2.yaml at fs=44100 (conf/exp/musicnet44k_4s.yaml): 8 octaves, 242 M parameters."""
    return NetConfig(num_octs=8, sample_rate=44100.0, audio_len=audio_len, Ns=[64, 64, 96, 96, 128, 128, 256, 256],
                     num_dils=[2, 3, 4, 5, 6, 7, 8, 8], attention_layers=[0, 0, 0, 0, 0, 1, 1, 1, 1], conv_mode=conv_mode)


def small_test(audio_len=16384, conv_mode=0):
    """A narrow network with the same topology (7 octaves, attention on the deep levels) for fast tests."""
    return NetConfig(audio_len=audio_len, Ns=[16, 16, 24, 24, 32, 32, 32], num_dils=[1, 2, 2, 3, 3, 3, 2],
                     attention_layers=[0, 0, 0, 0, 1, 1, 1, 1], conv_mode=conv_mode)
