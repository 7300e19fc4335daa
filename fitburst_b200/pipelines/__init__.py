"""Callers of the hot path: the fitburst_pipeline flow for one or many generic-format files."""
