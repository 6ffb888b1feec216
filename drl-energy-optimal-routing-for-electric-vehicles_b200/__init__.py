"""evrp_b200 — B200-native drop-in for the rollout hot path of
mengchengTang/DRL-Energy-optimal-Routing-for-Electric-Vehicles.

The package directory is named ``drl-energy-optimal-routing-for-electric-vehicles_b200`` (not importable as a
Python identifier); the repo-root shim ``evrp_b200/`` loads it under the name ``evrp_b200``.
"""
from . import _lib
from ._lib import EvrpLibraryError, lib
from .model import AttentionModel, AttentionModelAM
from .problem import (EVRPDataset, VehicleRoutingDataset, load_dataset, parse_cvrplib, save_dataset,
                      synthetic_instances, generate_instances)

__all__ = ["AttentionModel", "AttentionModelAM", "VehicleRoutingDataset", "parse_cvrplib", "synthetic_instances", "generate_instances", "save_dataset",
           "load_dataset", "EVRPDataset", "EvrpLibraryError", "lib"]
