"""sdfibm_b200 — B200-native (sm_100a) solid–fluid coupling path of sdfibm behind a C ABI."""
from . import capi  # noqa: F401

__all__ = ["capi"]
