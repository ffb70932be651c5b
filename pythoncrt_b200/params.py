"""Effect-parameter surface of the reference, kept name for name.

`CrtParams` carries the scalar arguments `process_video` forwards to the chain
(/root/reference/crt_filter.py:864-911); the three constructors accept what the
reference accepts:

  * `CrtParams.from_cli(argv)`   the argparse flags (crt_filter.py:1155-1206)
                                 with main()'s clamps (:1225-1260);
  * `CrtParams.from_preset(d)`   the GUI preset JSON keys written by
                                 `_collect_settings` (:2043-2080); unknown and
                                 encoder-only keys are ignored like
                                 `_apply_settings` does (:2093-2161);
  * keyword construction         process_video's own kwarg names.
"""
from __future__ import annotations

import argparse
import json
from dataclasses import asdict, dataclass, fields, replace
from typing import Any, Dict, Iterable, Optional

# preset key (crt_filter.py:2044-2080) -> CrtParams field
PRESET_KEYS = {
    "scanline": "scanline_strength",
    "triad": "triad_strength",
    "triad_gamma": "triad_gamma",
    "triad_softness": "triad_softness",
    "triad_preserve_luma": "triad_preserve_luma",
    "pixel_size": "pixel_size",
    "aberration_px": "aberration_px",
    "noise": "noise_strength",
    "bloom_sigma": "bloom_sigma",
    "bloom_strength": "bloom_strength",
    "bloom_threshold": "bloom_threshold",
    "vignette": "vignette_strength",
    "persistence": "persistence",
    "scanline_speed": "scanline_speed_px_s",
    "scanline_period": "scanline_period_px",
    "glitch_amp": "glitch_amp_px",
    "glitch_height": "glitch_height_frac",
    "fast_bloom": "fast_bloom",
    "brightness": "brightness",
    "contrast": "contrast",
    "gamma": "gamma",
    "saturation": "saturation",
    "temperature": "temperature",
    "flicker_strength": "flicker_strength",
    "flicker_hz": "flicker_hz",
    "grain_size": "grain_size",
    "scanline_angle": "scanline_angle",
    "scanline_thickness": "scanline_thickness",
    "warp_strength": "warp_strength",
}
# keys a preset file also carries that do not belong to the effect chain (:2062-2067)
PRESET_IGNORED = ("crf", "bitrate_kbps", "nvenc_preset", "gpu", "encoder")


@dataclass
class CrtParams:
    """Defaults are the CLI defaults (crt_filter.py:1160-1205)."""
    scanline_strength: float = 0.6
    triad_strength: float = 0.35
    triad_gamma: float = 2.2
    triad_preserve_luma: bool = False
    triad_softness: float = 0.5
    aberration_px: int = 1
    bloom_sigma: float = 1.2
    bloom_strength: float = 0.25
    bloom_threshold: float = 0.0
    noise_strength: float = 1.5
    vignette_strength: float = 0.25
    persistence: float = 0.2
    scanline_speed_px_s: float = 30.0
    scanline_period_px: float = 2.0
    fast_bloom: bool = True
    pixel_size: int = 2
    glitch_amp_px: int = 0
    glitch_height_frac: float = 0.0
    brightness: float = 0.0
    contrast: float = 1.0
    gamma: float = 1.0
    saturation: float = 1.0
    temperature: float = 0.0
    flicker_strength: float = 0.0
    flicker_hz: float = 0.0
    grain_size: int = 1
    scanline_angle: float = 0.0
    scanline_thickness: float = 1.0
    warp_strength: float = 0.0

    def but(self, **kw) -> "CrtParams":
        return replace(self, **kw)

    # ------------------------------------------------------------ GUI side --
    @classmethod
    def gui_defaults(cls) -> "CrtParams":
        """Widget defaults where they differ from the CLI (crt_filter.py:1463, :1493)."""
        return cls(triad_preserve_luma=True, scanline_speed_px_s=60.0)

    @classmethod
    def from_preset(cls, data: Dict[str, Any], base: Optional["CrtParams"] = None) -> "CrtParams":
        """Every key is optional; unknown keys are ignored (crt_filter.py:2093-2161)."""
        out = base if base is not None else cls.gui_defaults()
        if not isinstance(data, dict):
            return out
        types = {f.name: f.type for f in fields(cls)}
        upd = {}
        for key, name in PRESET_KEYS.items():
            if key in data:
                t = types[name]
                v = data[key]
                upd[name] = bool(v) if t == "bool" else (int(v) if t == "int" else float(v))
        return replace(out, **upd)

    @classmethod
    def load_preset(cls, path: str, base: Optional["CrtParams"] = None) -> "CrtParams":
        with open(path, "r", encoding="utf-8") as f:
            return cls.from_preset(json.load(f), base)

    def to_preset(self) -> Dict[str, Any]:
        d = asdict(self)
        return {key: d[name] for key, name in PRESET_KEYS.items()}

    # ------------------------------------------------------------ CLI side --
    @staticmethod
    def cli_parser() -> argparse.ArgumentParser:
        """The effect flags of parse_args (crt_filter.py:1160-1205): same names, types and defaults, from CLI_SURFACE."""
        p = argparse.ArgumentParser(add_help=False)
        for field, flag, kind, lo, hi in CLI_SURFACE:
            default = getattr(CrtParams, field)
            if kind is bool:
                if field == "fast_bloom":                 # a --flag / --no-flag pair defaulting to on (:1176-1178)
                    p.add_argument(flag, dest=field, action="store_true")
                    p.add_argument("--no-" + flag[2:], dest=field, action="store_false")
                    p.set_defaults(**{field: default})
                else:
                    p.add_argument(flag, dest=field, action="store_true")
            else:
                p.add_argument(flag, dest=field, type=kind, default=default)
        return p

    @classmethod
    def from_cli(cls, argv: Optional[Iterable[str]] = None) -> "CrtParams":
        """Parse the reference's flags (unknown flags such as --input are left
        alone) and apply main()'s clamps (crt_filter.py:1225-1260)."""
        a, _ = cls.cli_parser().parse_known_args(list(argv) if argv is not None else None)
        out = {}
        for field, _flag, kind, lo, hi in CLI_SURFACE:
            v = kind(getattr(a, field))
            if lo is not None:
                v = max(kind(lo), v)
            if hi is not None:
                v = min(kind(hi), v)
            out[field] = v
        return cls(**out)

    # ------------------------------------------------------- frame scalars --
    def phase_px(self, frame_index: int, fps: float) -> float:
        """scanline phase of frame i (crt_filter.py:1043)."""
        return (frame_index / float(fps)) * self.scanline_speed_px_s

    @staticmethod
    def time_sec(frame_index: int, fps: float) -> float:
        """flicker time of frame i (crt_filter.py:1064)."""
        return frame_index / float(fps)


# The reference's CLI surface for the effect chain as data: (CrtParams field, flag (:1160-1205), type, clamp low, clamp high
# (:1225-1260; None = unclamped)).  Defaults are the dataclass defaults.  Drives cli_parser() and from_cli().
CLI_SURFACE = (
    ("scanline_strength", "--scanline-strength", float, 0.0, 1.0),
    ("triad_strength", "--triad-strength", float, 0.0, 1.0),
    ("triad_gamma", "--triad-gamma", float, 0.1, None),
    ("triad_preserve_luma", "--triad-preserve-luma", bool, None, None),
    ("triad_softness", "--triad-softness", float, 0.0, None),
    ("aberration_px", "--aberration-px", int, -8, 8),
    ("bloom_sigma", "--bloom-sigma", float, 0.0, None),
    ("bloom_strength", "--bloom-strength", float, 0.0, None),
    ("bloom_threshold", "--bloom-threshold", float, 0.0, 1.0),
    ("noise_strength", "--noise-strength", float, 0.0, None),
    ("vignette_strength", "--vignette-strength", float, 0.0, 1.0),
    ("persistence", "--persistence", float, 0.0, 0.95),
    ("scanline_speed_px_s", "--scanline-speed", float, None, None),
    ("scanline_period_px", "--scanline-period", float, 1.0, None),
    ("fast_bloom", "--fast-bloom", bool, None, None),
    ("pixel_size", "--pixel-size", int, 1, None),
    ("glitch_amp_px", "--glitch-amp", int, 0, None),
    ("glitch_height_frac", "--glitch-height", float, 0.0, 1.0),
    ("brightness", "--brightness", float, None, None),
    ("contrast", "--contrast", float, None, None),
    ("gamma", "--gamma", float, 1e-3, None),
    ("saturation", "--saturation", float, 0.0, None),
    ("temperature", "--temperature", float, -1.0, 1.0),
    ("flicker_strength", "--flicker-strength", float, 0.0, 1.0),
    ("flicker_hz", "--flicker-hz", float, 0.0, None),
    ("grain_size", "--grain-size", int, 1, None),
    ("scanline_angle", "--scanline-angle", float, None, None),
    ("scanline_thickness", "--scanline-thickness", float, 0.1, None),
    ("warp_strength", "--warp-strength", float, -1.0, 1.0),
)
