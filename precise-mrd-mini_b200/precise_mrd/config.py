"""Configuration dataclasses -- same fields, same ``config_hash()`` as the reference
(``src/precise_mrd/config.py:13-94``), so hashes stamped into tables and JSON artifacts match
(``0b42f9b03f3fd555`` for configs/smoke.yaml at seed 7: SURVEY.md Appendix B).

One deliberate difference: ``load_config`` ignores YAML keys the
dataclasses do not know, so ``configs/default.yaml`` (which carries ``max_edit_distance``,
``min_alt_umi``, ``error_model``, ``filters``, ``qc``, ``output``) loads instead of raising.
"""
from __future__ import annotations

import dataclasses
import hashlib
import json
from dataclasses import asdict, dataclass
from pathlib import Path
from typing import Any, Dict, List, Optional

import yaml


@dataclass
class SimulationConfig:
    allele_fractions: List[float]
    umi_depths: List[int]
    n_replicates: int
    n_bootstrap: int


@dataclass
class UMIConfig:
    min_family_size: int
    max_family_size: int
    quality_threshold: int
    consensus_threshold: float


@dataclass
class StatsConfig:
    test_type: str
    alpha: float
    fdr_method: str


@dataclass
class LODConfig:
    detection_threshold: float
    confidence_level: float


@dataclass
class FASTQConfig:
    input_path: str
    max_reads: Optional[int] = None
    umi_pattern: Optional[str] = None
    quality_threshold: int = 20
    min_family_size: int = 3


@dataclass
class PipelineConfig:
    run_id: str
    seed: int
    umi: UMIConfig
    stats: StatsConfig
    lod: LODConfig
    simulation: Optional[SimulationConfig] = None
    fastq: Optional[FASTQConfig] = None

    def to_dict(self) -> Dict[str, Any]:
        return asdict(self)

    def config_hash(self) -> str:
        config_str = json.dumps(self.to_dict(), sort_keys=True)
        return hashlib.sha256(config_str.encode()).hexdigest()[:16]


def _known(cls, data: Dict[str, Any]):
    names = {f.name for f in dataclasses.fields(cls)}
    return cls(**{k: v for k, v in data.items() if k in names})


def load_config(path: str | Path) -> PipelineConfig:
    """YAML -> PipelineConfig (``config.py:77-89``), tolerant of unknown keys."""
    with open(path, "r") as f:
        data = yaml.safe_load(f)
    return PipelineConfig(
        run_id=data["run_id"],
        seed=data["seed"],
        simulation=_known(SimulationConfig, data["simulation"]) if data.get("simulation") else None,
        umi=_known(UMIConfig, data["umi"]),
        stats=_known(StatsConfig, data["stats"]),
        lod=_known(LODConfig, data["lod"]),
        fastq=_known(FASTQConfig, data["fastq"]) if data.get("fastq") else None,
    )


def dump_config(config: PipelineConfig, path: str | Path) -> None:
    with open(path, "w") as f:
        yaml.safe_dump(config.to_dict(), f, default_flow_style=False)


def section(config: Any, name: str) -> Dict[str, Any]:
    """Stages accept dataclass OR dict sub-sections (``simulate.py:38-52``, ``collapse.py:35-44``,
    ``call.py:149-156``; exercised by ``tests/test_rng_api.py:151-172``)."""
    sec = getattr(config, name) if not isinstance(config, dict) else config[name]
    if sec is None:
        raise ValueError(f"config.{name} is required")
    return sec if isinstance(sec, dict) else asdict(sec)
