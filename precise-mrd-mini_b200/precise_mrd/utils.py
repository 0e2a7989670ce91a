"""``PipelineIO`` (``src/precise_mrd/utils.py:11-31``)."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict


class PipelineIO:
    @staticmethod
    def ensure_dir(path) -> Path:
        p = Path(path)
        p.mkdir(parents=True, exist_ok=True)
        return p

    @staticmethod
    def save_json(data: Dict[str, Any], path) -> None:
        PipelineIO.ensure_dir(Path(path).parent)
        with open(path, "w") as f:
            json.dump(data, f, indent=2, default=str)

    @staticmethod
    def load_json(path) -> Dict[str, Any]:
        with open(path, "r") as f:
            return json.load(f)


def json_ready(obj: Any) -> Any:
    """numpy scalars -> Python scalars before ``json.dump`` (``cli.py:67-85``)."""
    import numpy as np

    if isinstance(obj, dict):
        return {k: json_ready(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [json_ready(v) for v in obj]
    if isinstance(obj, np.ndarray):
        return [json_ready(v) for v in obj.tolist()]
    if isinstance(obj, (np.bool_,)):
        return bool(obj)
    if isinstance(obj, (np.floating, float)):
        return float(obj)
    if isinstance(obj, (np.integer, int)) and not isinstance(obj, bool):
        return int(obj)
    return obj
