"""ORACLE — test infrastructure only (see oracle/bm25_oracle.h). Never imported by diagon_b200/."""
