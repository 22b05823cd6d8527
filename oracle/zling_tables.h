/* Format-defining constants of the zling bitstream (values are facts of the format, not code).
 * TEST-INFRASTRUCTURE copy used by oracle/zling_oracle.c; the product has its own copy in
 * libzling_b200/csrc/zl_tables.h.  Provenance (reference generator src/tables/gen.py):
 *   - initial MTF order: the 256-entry permutation written at gen.py:31-49 (hex string below);
 *   - mtfnext[i] = floor(0.95 i) for i < 128, floor(0.55 i) otherwise        (gen.py:51-56);
 *   - match-index buckets: extra bits 0,0,0,0,1,1,2,2,...,7,7 then 8 for the rest; base/code derived (gen.py:10-19).
 * zo_tables_init() expands the rules; tests/test_tables.py checks every value against the reference .inc files.
 */
#ifndef ZLING_ORACLE_TABLES_H
#define ZLING_ORACLE_TABLES_H
static const char zo_mtfinit_hex[513] =
    "20657461696f6e72736c686463755d5b6d7067660a796227772e2c763b267c2f"
    "316b3d3043413a2d54533c3e327149392a782928424d50454435334846383447"
    "52364c374e577a7d7b4f6a554a4bd05fc32356d75a2259d180e0b8835ce32521"
    "b0a9cee2823f5851a1992b81bcb3d8a4b5bd94beadbbbae5e1a7d9b1b2a895b9"
    "c59093c4cfc2b49c84aaa688b6bf09e68da0af24988ca5915e85a3b7ab9d89ae"
    "8687ec97e79bc99e8a8f96a29f8bac9a7ee8eb92e9e4cacb8ed6edccdbead560"
    "dac740d2efc6d3cdd4f0dedcc80001020304050607080b0c0d0e0f1011121314"
    "15161718191a1b1c1d1e1f7fc0c1dddfeef1f2f3f4f5f6f7f8f9fafbfcfdfeff";
#endif
