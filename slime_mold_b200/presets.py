"""Presets -- mirror of /root/reference/src/presets.rs (Preset, PresetManager, the 8 shipped presets)."""
from __future__ import annotations

from typing import List, Optional

from .settings import Settings


class Preset:
    """presets.rs:6-15"""

    def __init__(self, name: str, settings: Settings):
        self.name = name
        self.settings = settings


class PresetManager:
    """presets.rs:17-37"""

    def __init__(self) -> None:
        self.presets: List[Preset] = []

    def add_preset(self, preset: Preset) -> None:
        self.presets.append(preset)

    def get_preset(self, name: str) -> Optional[Preset]:
        for p in self.presets:
            if p.name == name:
                return p
        return None

    def get_preset_names(self) -> List[str]:
        return [p.name for p in self.presets]


def init_preset_manager() -> PresetManager:
    """presets.rs:45-155: Default, Sponge, Firecracker Trees, Threads, Curls, Waves, Snake, Mesh."""
    pm = PresetManager()
    d = Settings.default()
    pm.add_preset(Preset("Default", d.clone()))
    pm.add_preset(Preset("Sponge", d.clone(                                    # presets.rs:48-62
        agent_jitter=0.0, agent_speed_min=20.0, agent_speed_max=30.0, agent_turn_speed=0.43,
        agent_sensor_angle=0.3, agent_sensor_distance=20.0, pheromone_deposition_amount=1.0,
        pheromone_decay_factor=1.0, pheromone_diffusion_rate=1.0)))
    pm.add_preset(Preset("Firecracker Trees", d.clone(                         # presets.rs:63-77
        agent_jitter=0.1, agent_speed_min=60.0, agent_speed_max=60.0, agent_turn_speed=1.47,
        agent_sensor_angle=0.3, agent_sensor_distance=20.0, pheromone_deposition_amount=1.0,
        pheromone_decay_factor=10.0, pheromone_diffusion_rate=1.0)))
    pm.add_preset(Preset("Threads", d.clone(                                   # presets.rs:78-92
        agent_jitter=0.0, agent_speed_min=70.0, agent_speed_max=80.0, agent_turn_speed=0.02,
        agent_sensor_angle=0.3, agent_sensor_distance=20.0, pheromone_deposition_amount=1.0,
        pheromone_decay_factor=10.0, pheromone_diffusion_rate=0.1)))
    pm.add_preset(Preset("Curls", d.clone(                                     # presets.rs:93-108
        agent_count=3_000_000, agent_jitter=5.0, agent_speed_min=70.0, agent_speed_max=80.0,
        agent_turn_speed=0.05, agent_sensor_angle=0.3, agent_sensor_distance=20.0,
        pheromone_deposition_amount=1.0, pheromone_decay_factor=75.0, pheromone_diffusion_rate=0.1)))
    pm.add_preset(Preset("Waves", d.clone(                                     # presets.rs:109-123
        agent_jitter=1.0, agent_speed_min=30.0, agent_speed_max=50.0, agent_turn_speed=6.0,
        agent_sensor_angle=0.3, agent_sensor_distance=20.0, pheromone_deposition_amount=1.0,
        pheromone_decay_factor=10.0, pheromone_diffusion_rate=0.1)))
    pm.add_preset(Preset("Snake", d.clone(                                     # presets.rs:124-138
        agent_jitter=3.0, agent_speed_min=100.0, agent_speed_max=120.0, agent_turn_speed=0.37,
        agent_sensor_angle=1.34, agent_sensor_distance=225.0, pheromone_deposition_amount=1.0,
        pheromone_decay_factor=10.0, pheromone_diffusion_rate=1.0)))
    pm.add_preset(Preset("Mesh", d.clone(                                      # presets.rs:139-153
        agent_jitter=3.0, agent_speed_min=100.0, agent_speed_max=120.0, agent_turn_speed=6.0,
        agent_sensor_angle=1.57, agent_sensor_distance=225.0, pheromone_deposition_amount=1.0,
        pheromone_decay_factor=10.0, pheromone_diffusion_rate=1.0)))
    return pm
