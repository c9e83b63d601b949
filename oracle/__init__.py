"""Parity oracle for the MeShClust2 hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (meshclust2_b200/) never does.

  oracle.port  -> ctypes view of oracle/libmc2oracle.so  (plain-C restatement, mc2_oracle.c)
  oracle.ref   -> ctypes view of oracle/_ref/libmc2ref.so (the unmodified reference compiled from
                  /root/reference by oracle/Makefile; present only where it was built)
"""
