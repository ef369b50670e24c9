"""Consumers of the neighbor list that this package fuses with the stencil sweep."""
from . import electrostatics  # noqa: F401
