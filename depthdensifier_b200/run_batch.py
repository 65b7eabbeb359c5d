"""Batch driver (reference: scripts/run_batch.py:26-110): run the densification pipeline on every scan folder
under ``root_dir`` (each holding ``sparse/0`` and ``images``), keep going when a scan fails, print a timing
table at the end.  Same ``BatchConfig`` fields and flags: ``root_dir`` and ``output_dir`` are required, ``config``
embeds the ScriptConfig shared by all scans (e.g. ``--config.filtering.vote-threshold 3``), and each scan's model
goes to ``<output_dir>/<scan>/sparse/0`` (scripts/run_batch.py:63-65)."""

from __future__ import annotations

import copy
import time
from dataclasses import dataclass, field
from pathlib import Path

from .pipeline import PathsConfig, ScriptConfig, main as densify_main


@dataclass
class BatchConfig:
    """Batch run over every scan folder of one directory."""

    root_dir: Path
    """Folder whose sub-folders are the scans (each with images/ and sparse/0/)."""
    output_dir: Path
    """Results go to <output_dir>/<scan>/sparse/0."""
    config: ScriptConfig = field(default_factory=ScriptConfig)
    """Settings shared by all scans; the three paths are replaced per scan."""
    depth_root: Path | None = None
    """new: <depth_root>/<scan> holds the precomputed depth .npz files (None -> MoGe)."""


def main(config: BatchConfig) -> dict[str, float | str]:
    root = Path(config.root_dir).resolve()
    if not root.is_dir():  # scripts/run_batch.py:48-50
        print(f"Error: Root directory not found at {root}")
        return {}
    scans = sorted(p for p in root.iterdir() if p.is_dir())
    if not scans:
        print(f"No scan folders found in {config.root_dir}")
        return {}
    report: dict[str, float | str] = {}
    t_all = time.time()
    for scan in scans:
        recon, images = scan / "sparse" / "0", scan / "images"
        if not recon.is_dir() or not images.is_dir():
            print(f"Skipping {scan.name}: missing sparse/0 or images")
            continue
        run_config = copy.deepcopy(config.config)
        run_config.paths = PathsConfig(recon_path=recon, image_dir=images, output_model_dir=Path(config.output_dir) / scan.name / "sparse" / "0",
                                       depth_dir=(Path(config.depth_root) / scan.name) if config.depth_root is not None else None)
        t0 = time.time()
        try:
            densify_main(run_config)
            report[scan.name] = time.time() - t0
        except Exception as e:  # one bad scan must not stop the batch (scripts/run_batch.py:82-91)
            print(f"!!!!!! FAILED to process {scan.name}: {e}")
            report[scan.name] = "FAILED"
    print("\n" + "=" * 50 + "\nBatch processing report\n" + "=" * 50)
    for name, val in report.items():
        print(f"{name:<30} {val if isinstance(val, str) else f'{val:.2f}s'}")
    print(f"{'total':<30} {time.time() - t_all:.2f}s")
    return report


def cli() -> None:
    import tyro  # noqa: PLC0415

    main(tyro.cli(BatchConfig))


if __name__ == "__main__":
    cli()
