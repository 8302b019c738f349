"""Seeded synthetic inputs of SURVEY.md section 8(d): re-export of the repo-level ``workloads`` module (the benches import that one
directly, so that their timed arm touches nothing under oracle/)."""
from workloads import _rescale, config, macs, make_scene, make_template  # noqa: F401
