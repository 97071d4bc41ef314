"""Replay side of the hot path (reference: slimdqn/sample_collection/__init__.py:3)."""
from typing import NewType

ReplayItemID = NewType("ReplayItemID", int)
