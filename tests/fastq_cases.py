"""FASTQ grammar cases (FastqReader.hpp:42-110).  Shared by the oracle tests
and the host-reader tests of the product."""

GOOD = {
    "plain": b"@r1\nACGT\n+\nIIII\n@r2\nGGTTA\n+\nIIIII\n",
    "no_final_newline": b"@r1\nACGT\n+\nIIII\n@r2\nGGTTA\n+\nIIIII",
    "empty_file": b"",
    "name_with_spaces": b"@r1 extra stuff\tmore\nACGT\n+r1 again\nIIII\n",
    "multiline_seq": b"@r1\nAC\nGT\nTT\n+\nIIIIII\n@r2\nA\n+\nI\n",
    "multiline_qual": b"@r1\nACGTAC\n+\nIII\nIII\n@r2\nAA\n+\nII\n",
    "crlf": b"@r1\r\nACGT\r\n+\r\nIIII\r\n@r2\r\nGG\r\n+\r\nII\r\n",
    "qual_starts_with_at": b"@r1\nACGT\n+\n@III\n@r2\nGG\n+\n@@\n",
    "empty_seq": b"@r1\n\n+\n\n@r2\nAC\n+\nII\n",
    "lower_and_n": b"@r1\nacgtNNRY.x\n+\nIIIIIIIIII\n",
    "long_read": b"@r1\n" + b"ACGT" * 200 + b"\n+\n" + b"I" * 800 + b"\n",
}

BAD = {
    "no_at": (b"r1\nACGT\n+\nIIII\n", "read name should start with '@' (starting line 1)"),
    "no_at_second": (b"@r1\nACGT\n+\nIIII\nr2\nAC\n+\nII\n", "read name should start with '@' (starting line 5)"),
    "truncated_name": (b"@r1", "premature end of the file at line 1"),
    "truncated_seq": (b"@r1\nACGT", "premature end of the file at line 2"),
    "truncated_plus": (b"@r1\nACGT\n+", "premature end of the file at line 3"),
    "short_qual": (b"@r1\nACGT\n+\nIII\n", "non-equal lengths for quality and sequence strings (starting line 1)"),
    "short_qual_eof": (b"@r1\nACGT\n+\nIII", "non-equal lengths for quality and sequence strings (starting line 1)"),
    "qual_longer_multiline": (b"@r1\nACGT\n+\nII\nIII\n", "non-equal lengths for quality and sequence strings (starting line 1)"),
    "long_qual": (b"@r1\nACGT\n+\nIIIII\n", "non-equal lengths for quality and sequence strings (starting line 1)"),
    "second_record_short": (b"@r1\nACGT\n+\nIIII\n@r2\nACGT\n+\nII\n", "non-equal lengths for quality and sequence strings (starting line 5)"),
}
