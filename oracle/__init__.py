"""CPU oracle for the Polmux/Optilux SSFM fiber path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``polmux_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` do, and there only as the checker
or as the timed CPU baseline.

Parity status: see the header of ``oracle/fiber_oracle.py`` and DESIGN.md.
"""
