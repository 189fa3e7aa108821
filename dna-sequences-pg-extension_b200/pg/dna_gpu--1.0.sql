-- Additions to dna--1.0.sql for the GPU build.  generate_kmers keeps its declaration
-- (dna--1.0.sql:188-191); only the implementation behind MODULE_PATHNAME changes.

-- total / distinct / unique of README.md:122-130 computed on the GPU in one call
CREATE OR REPLACE FUNCTION kmer_stats(dna, integer, OUT total bigint, OUT "distinct" bigint, OUT uniq bigint)
    RETURNS record
    AS 'MODULE_PATHNAME', 'kmer_stats'
    LANGUAGE C IMMUTABLE STRICT PARALLEL SAFE;

-- SELECT kmer, count(*) FROM generate_kmers(seq, k) GROUP BY kmer  (README.md:107-116) in one call
CREATE OR REPLACE FUNCTION count_kmers(dna, integer, OUT kmer kmer, OUT count bigint)
    RETURNS SETOF record
    AS 'MODULE_PATHNAME', 'count_kmers'
    LANGUAGE C IMMUTABLE STRICT PARALLEL SAFE;
