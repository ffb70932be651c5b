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
        """The effect flags of parse_args (crt_filter.py:1160-1205), same names and defaults."""
        p = argparse.ArgumentParser(add_help=False)
        p.add_argument("--scanline-strength", type=float, default=0.6)
        p.add_argument("--triad-strength", type=float, default=0.35)
        p.add_argument("--triad-gamma", type=float, default=2.2)
        p.add_argument("--triad-preserve-luma", action="store_true")
        p.add_argument("--triad-softness", type=float, default=0.5)
        p.add_argument("--aberration-px", type=int, default=1)
        p.add_argument("--bloom-sigma", type=float, default=1.2)
        p.add_argument("--bloom-strength", type=float, default=0.25)
        p.add_argument("--bloom-threshold", type=float, default=0.0)
        p.add_argument("--noise-strength", type=float, default=1.5)
        p.add_argument("--vignette-strength", type=float, default=0.25)
        p.add_argument("--persistence", type=float, default=0.2)
        p.add_argument("--scanline-speed", type=float, default=30.0)
        p.add_argument("--scanline-period", type=float, default=2.0)
        p.add_argument("--fast-bloom", action="store_true")
        p.add_argument("--no-fast-bloom", dest="fast_bloom", action="store_false")
        p.set_defaults(fast_bloom=True)
        p.add_argument("--pixel-size", type=int, default=2)
        p.add_argument("--brightness", type=float, default=0.0)
        p.add_argument("--contrast", type=float, default=1.0)
        p.add_argument("--gamma", type=float, default=1.0)
        p.add_argument("--saturation", type=float, default=1.0)
        p.add_argument("--temperature", type=float, default=0.0)
        p.add_argument("--flicker-strength", type=float, default=0.0)
        p.add_argument("--flicker-hz", type=float, default=0.0)
        p.add_argument("--grain-size", type=int, default=1)
        p.add_argument("--scanline-angle", type=float, default=0.0)
        p.add_argument("--scanline-thickness", type=float, default=1.0)
        p.add_argument("--warp-strength", type=float, default=0.0)
        p.add_argument("--glitch-amp", type=int, default=0)
        p.add_argument("--glitch-height", type=float, default=0.0)
        return p

    @classmethod
    def from_cli(cls, argv: Optional[Iterable[str]] = None) -> "CrtParams":
        """Parse the reference's flags (unknown flags such as --input are left
        alone) and apply main()'s clamps (crt_filter.py:1225-1260)."""
        a, _ = cls.cli_parser().parse_known_args(list(argv) if argv is not None else None)
        return cls(
            scanline_strength=float(max(0.0, min(1.0, a.scanline_strength))),
            triad_strength=float(max(0.0, min(1.0, a.triad_strength))),
            triad_gamma=float(max(0.1, a.triad_gamma)),
            triad_preserve_luma=bool(a.triad_preserve_luma),
            triad_softness=float(max(0.0, a.triad_softness)),
            aberration_px=int(max(-8, min(8, a.aberration_px))),
            bloom_sigma=max(0.0, a.bloom_sigma),
            bloom_strength=max(0.0, a.bloom_strength),
            noise_strength=max(0.0, a.noise_strength),
            vignette_strength=float(max(0.0, min(1.0, a.vignette_strength))),
            persistence=float(max(0.0, min(0.95, a.persistence))),
            scanline_speed_px_s=float(a.scanline_speed),
            scanline_period_px=max(1.0, float(a.scanline_period)),
            fast_bloom=bool(a.fast_bloom),
            pixel_size=max(1, int(a.pixel_size)),
            glitch_amp_px=max(0, int(a.glitch_amp)),
            glitch_height_frac=float(max(0.0, min(1.0, a.glitch_height))),
            bloom_threshold=float(max(0.0, min(1.0, a.bloom_threshold))),
            brightness=float(a.brightness),
            contrast=float(a.contrast),
            gamma=float(max(1e-3, a.gamma)),
            saturation=float(max(0.0, a.saturation)),
            temperature=float(max(-1.0, min(1.0, a.temperature))),
            flicker_strength=float(max(0.0, min(1.0, a.flicker_strength))),
            flicker_hz=float(max(0.0, a.flicker_hz)),
            grain_size=max(1, int(a.grain_size)),
            scanline_angle=float(a.scanline_angle),
            scanline_thickness=float(max(0.1, a.scanline_thickness)),
            warp_strength=float(max(-1.0, min(1.0, a.warp_strength))),
        )

    # ------------------------------------------------------- frame scalars --
    def phase_px(self, frame_index: int, fps: float) -> float:
        """scanline phase of frame i (crt_filter.py:1043)."""
        return (frame_index / float(fps)) * self.scanline_speed_px_s

    @staticmethod
    def time_sec(frame_index: int, fps: float) -> float:
        """flicker time of frame i (crt_filter.py:1064)."""
        return frame_index / float(fps)
