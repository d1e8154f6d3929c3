"""CPU oracle for the likelihood/score path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package (see oracle/oracle.h).  The product
package `apophenia_b200` never does.
"""
from .oracle import *  # noqa: F401,F403
