"""Settings -- mirror of /root/reference/src/settings.rs (same field names, same defaults).

The reference's `Settings` struct stays the drop-in configuration surface; the
backend consumes it through `SimSizeUniform.new(width, height, decay_factor, settings)`
exactly like src/main.rs:49-66.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field, replace
from typing import Tuple

# General settings (settings.rs:4-6)
DEFAULT_WIDTH = 1600
DEFAULT_HEIGHT = 900
DEFAULT_IS_FULLSCREEN = False

# Agent settings (settings.rs:9-17)
AGENT_COUNT = 10_000_000
AGENT_SPEED_MIN = 30.0
AGENT_SPEED_MAX = 50.0
AGENT_TURN_SPEED = 0.43
AGENT_POSSIBLE_STARTING_HEADINGS = (0.0, 360.0)
DEPOSITION_AMOUNT = 1.0
AGENT_JITTER = 0.0
AGENT_SENSOR_ANGLE = 0.3
AGENT_SENSOR_DISTANCE = 20.0

# Pheromone settings (settings.rs:21-27)
DECAY_FACTOR = 10.0
DIFFUSION_RATE = 1.0
BLUR_RADIUS = 2.0
BLUR_SIGMA = 1.0


@dataclass
class Settings:
    """settings.rs:29-47, field for field."""

    agent_count: int = AGENT_COUNT
    agent_jitter: float = AGENT_JITTER
    agent_possible_starting_headings: Tuple[float, float] = AGENT_POSSIBLE_STARTING_HEADINGS
    agent_speed_max: float = AGENT_SPEED_MAX
    agent_speed_min: float = AGENT_SPEED_MIN
    agent_turn_speed: float = AGENT_TURN_SPEED
    pheromone_decay_factor: float = DECAY_FACTOR
    pheromone_diffusion_rate: float = DIFFUSION_RATE
    pheromone_deposition_amount: float = DEPOSITION_AMOUNT
    window_fullscreen: bool = DEFAULT_IS_FULLSCREEN
    window_height: int = DEFAULT_HEIGHT
    window_width: int = DEFAULT_WIDTH
    agent_sensor_angle: float = AGENT_SENSOR_ANGLE
    agent_sensor_distance: float = AGENT_SENSOR_DISTANCE
    blur_radius: float = BLUR_RADIUS
    blur_sigma: float = BLUR_SIGMA

    @staticmethod
    def default() -> "Settings":
        return Settings()

    def clone(self, **changes) -> "Settings":
        return replace(self, **changes)


class SimSizeUniform(C.Structure):
    """`#[repr(C)] struct SimSizeUniform`, src/main.rs:29-46 -- 56 bytes; == C `sm_params`."""

    _fields_ = [
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("decay_factor", C.c_float),
        ("agent_jitter", C.c_float),
        ("agent_speed_min", C.c_float),
        ("agent_speed_max", C.c_float),
        ("agent_turn_speed", C.c_float),
        ("agent_sensor_angle", C.c_float),
        ("agent_sensor_distance", C.c_float),
        ("diffusion_rate", C.c_float),
        ("pheromone_deposition_amount", C.c_float),
        ("blur_radius", C.c_float),
        ("blur_sigma", C.c_float),
        ("_pad", C.c_uint32),
    ]

    @classmethod
    def new(cls, width: int, height: int, decay_factor: float, settings: Settings) -> "SimSizeUniform":
        """SimSizeUniform::new, src/main.rs:49-66."""
        return cls(
            width=width,
            height=height,
            decay_factor=decay_factor,
            agent_jitter=settings.agent_jitter,
            agent_speed_min=settings.agent_speed_min,
            agent_speed_max=settings.agent_speed_max,
            agent_turn_speed=settings.agent_turn_speed,
            agent_sensor_angle=settings.agent_sensor_angle,
            agent_sensor_distance=settings.agent_sensor_distance,
            diffusion_rate=settings.pheromone_diffusion_rate,
            pheromone_deposition_amount=settings.pheromone_deposition_amount,
            blur_radius=settings.blur_radius,
            blur_sigma=settings.blur_sigma,
            _pad=0,
        )

    def to_bytes(self) -> bytes:
        return bytes(self)


assert C.sizeof(SimSizeUniform) == 56
