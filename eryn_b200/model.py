"""Carrier handed to every move — same six fields as eryn.model.Model (model.py:8-18)."""
from collections import namedtuple

__all__ = ["Model"]

Model = namedtuple(
    "Model",
    ("log_like_fn", "compute_log_like_fn", "compute_log_prior_fn", "temperature_control", "map_fn", "random"),
)
