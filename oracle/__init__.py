"""TEST INFRASTRUCTURE -- CPU oracle for the batched LDPC syndrome decoder.

`oracle.cpu` wraps oracle/liboracle.so (our C restatement of the reference decoders);
`oracle.ref` wraps oracle/_ref/libqkdref.so (the UNMODIFIED reference sources compiled behind a C driver).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
package. The product (qkd_ldpc_v_b200) never does.
"""
