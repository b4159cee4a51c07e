"""ORACLE — CPU restatement of the reference's tokenization front end.

Test infrastructure only.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from the product package.
"""
