"""CPU oracle for the TrackNetV3 hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package, and only as the checker / CPU baseline. Nothing under tracknetv3_b200/ imports it.
"""
