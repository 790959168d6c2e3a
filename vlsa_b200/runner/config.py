"""Flat-YAML experiment configs of the reference, parsed unchanged.

Mirrors the semantics (not the code) of main.py:72-103 + utils/func.py:284-321 (`args_grid`: every
list-valued key spans a grid axis) and runner/base_handler.py:46-72 (`{0}` = dataset name, `{2}` = split
seed, null `vlsa_img_encoder_num_query` = number of prototype sentences).  Only what the hot path reads.
"""
from __future__ import annotations

import itertools
from typing import Any

import yaml

# prototypes per dataset (runner/global_cfg.py:1-21 of the reference: len(prototype texts))
NUM_QUERY = {"tcga_blca": 12, "tcga_brca": 10, "tcga_gbmlgg": 8, "tcga_luad": 7, "tcga_ucec": 8}


def load_config(path: str) -> dict:
    with open(path) as fh:
        return yaml.safe_load(fh)


def expand_grid(cfg: dict) -> list[dict]:
    """One resolved dict per point of the grid spanned by the list-valued keys (first key varies slowest)."""
    fixed = {k: v for k, v in cfg.items() if not isinstance(v, list)}
    axes = {k: v for k, v in cfg.items() if isinstance(v, list)}
    out = []
    for combo in itertools.product(*axes.values()):
        d = dict(fixed)
        d.update(dict(zip(axes.keys(), combo)))
        out.append(d)
    return out


def _fill(value: Any, fill: Any, ind: str):
    if isinstance(value, str) and ind in value:
        new = value.replace(ind, str(fill))
        return new
    return value


def resolve_placeholders(cfg: dict, num_query: int | None = None, time_bins: int | None = None) -> dict:
    """Fill `{0}` / `{2}` placeholders and a null `vlsa_img_encoder_num_query`; set the number of ranks."""
    cfg = dict(cfg)
    name = cfg.get("dataset_name", "")
    for k in ("path_patch", "path_coord", "path_cluster", "path_graph", "path_table", "data_split_path",
              "vlsa_img_encoder_query_text_load_idx"):
        if k in cfg:
            cfg[k] = _fill(cfg[k], name, "{0}")
    if "save_path" in cfg:
        cfg["save_path"] = _fill(cfg["save_path"], name[5:], "{0}")
    if "data_split_path" in cfg:
        cfg["data_split_path"] = _fill(cfg["data_split_path"], cfg.get("data_split_seed", 0), "{2}")
    key = "vlsa_img_encoder_num_query"
    if key in cfg and cfg[key] is None:
        cfg[key] = int(num_query if num_query is not None else NUM_QUERY.get(name, 12))
    if time_bins is not None:
        cfg["time_bins"] = int(time_bins)
        for k in ("vlsa_pmt_learner_coop_num_ranks", "vlsa_pmt_learner_adapter_num_ranks"):
            if k in cfg:
                cfg[k] = int(time_bins)
    return cfg
