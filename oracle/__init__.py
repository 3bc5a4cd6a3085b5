"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header)."""
from .oracle import *  # noqa: F401,F403
