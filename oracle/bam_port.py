"""TEST INFRASTRUCTURE ONLY (see oracle/pyoracle.py): CPU restatement of BAM_handler::get_reads
(/root/reference/pepper_variant/modules/cpp/bam_handler.cpp:115-444) over already-decoded BAM records, written as the
same per-base loops as the reference. PINNED (round 2): tests/test_ingest.py::test_get_reads_matches_compiled_reference
checks it against the UNMODIFIED bam_handler.cpp compiled over oracle/hts_mini (oracle/_ref/pv_ref_bam) and against the
digests that reference wrote (tests/golden/bam_get_reads.json, tests/golden/make_bam_golden.py).

Iterator contract of htslib sam_itr_queryi(idx, tid, beg, end) as used at :128: records of `tid` with
pos < end and bam_endpos > beg, in file order."""


def hp_tag(aux):
    """HP integer tag, walking the aux block like bam_handler.cpp:313-421."""
    import struct
    s, end, hp = 0, len(aux), 0
    size = {"A": 1, "c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}
    fmt = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}
    while end - s >= 4:
        tag = aux[s:s + 2]; t = chr(aux[s + 2]); s += 3
        if t == "A":
            s += 1
        elif t in fmt:
            if end - s < size[t]:
                return hp
            v, = struct.unpack_from(fmt[t], aux, s)
            if tag == b"HP":
                hp = v
            s += size[t]
        elif t == "f":
            if end - s < 4:
                return hp
            s += 4
        elif t in "ZH":
            while s < end and aux[s]:
                s += 1
            if s >= end:
                return hp
            s += 1
        elif t == "B":
            if end - s < 5:
                return hp
            st = chr(aux[s])
            if st not in size:
                return hp
            n, = struct.unpack_from("<I", aux, s + 1)
            s += 5 + n * size[st]
        else:
            return hp
    return hp


def get_reads(records, tid, start, stop, include_supplementary, min_mapq=0, min_baseq=0):
    """-> list of dicts(pos, pos_end, sequence, base_qualities, cigar_tuples [(op, len)], is_reverse, mapping_quality,
    hp_tag, query_name)."""
    out = []
    for rec in records:
        if rec["tid"] != tid:
            continue
        rlen = sum(l for op, l in rec["cigar"] if op in (0, 2, 3, 7, 8))
        endpos = rec["pos"] + (1 if (rec["flag"] & 4) or rlen == 0 else rlen)
        if not (rec["pos"] < stop and endpos > start):
            continue
        flag = rec["flag"]
        if flag & 0x200 or flag & 0x400 or flag & 0x100 or flag & 0x4:          # :133-136
            continue
        if not include_supplementary and flag & 0x800:                          # :137-139
            continue
        if rec["mapq"] < min_mapq:                                              # :142-144
            continue
        seq, qual = rec["seq"], rec["qual"]
        read_seq, quals, tuples = [], [], []
        pos_start = pos_end = -1
        cur_pos, cur_idx = rec["pos"], 0
        bad = False
        for op, clen in rec["cigar"]:
            if cur_pos > stop:                                                  # :180-182
                break
            mod = 0
            if op in (0, 7, 8):                                                 # :185-225
                ci = 0
                if cur_pos < start:
                    ci = min(start - cur_pos, clen)
                    cur_idx += ci; cur_pos += ci
                for _ in range(ci, clen):
                    if cur_pos <= stop:
                        if pos_start == -1:
                            pos_start = cur_pos; pos_end = pos_start
                        if cur_idx >= len(seq):
                            bad = True; break
                        quals.append(qual[cur_idx]); read_seq.append(seq[cur_idx].upper())
                        mod += 1; pos_end += 1
                    else:
                        break
                    cur_idx += 1; cur_pos += 1
            elif op in (4, 1):                                                  # :226-262
                if start <= cur_pos <= stop and pos_start != -1:
                    for _ in range(clen):
                        if cur_idx >= len(seq):
                            bad = True; break
                        quals.append(qual[cur_idx]); read_seq.append(seq[cur_idx].upper())
                        mod += 1; cur_idx += 1
                else:
                    cur_idx += clen
            elif op in (3, 2):                                                  # :263-291
                if start <= cur_pos <= stop and pos_start != -1:
                    for _ in range(clen):
                        if cur_pos <= stop:
                            mod += 1; pos_end += 1
                        else:
                            break
                        cur_pos += 1
                else:
                    cur_pos += clen
            if bad:
                break
            if mod > 0:
                tuples.append((op, mod))
        if bad or not read_seq:                                                 # :424
            continue
        out.append(dict(pos=pos_start, pos_end=pos_end, sequence="".join(read_seq), base_qualities=quals,
                        cigar_tuples=tuples, is_reverse=bool(flag & 0x10), mapping_quality=rec["mapq"],
                        hp_tag=hp_tag(rec["aux"]), query_name=rec["name"]))
    return out
