"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's joint + transducer-loss path.  Nothing under
``transformer-transducer_b200/`` or ``warprnnt_pytorch/`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do, and only as the checker.
"""
