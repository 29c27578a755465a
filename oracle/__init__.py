"""CPU oracle for the TFSWA-UNet hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tfswa-unet_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker
or as the timed CPU baseline - never as the product path.
"""
