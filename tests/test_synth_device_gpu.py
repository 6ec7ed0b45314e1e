"""The device-side synthetic generator (csrc/synth_device.cu) is the bit-identical twin of the seeded host generator
(csrc/synth_reads.c): same reads, bases, qualities, CIGARs, reference for the same (profile, seed, regions)."""
import numpy as np
import pytest
import torch

from pepper_thesis_b200 import device as dev, synth, synth_device

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("profile,cov,first,n", [("ont_r9", 12.0, 0, 3), ("ont_r10", 9.0, 5, 2), ("hifi", 8.0, 2, 2)])
def test_device_generator_equals_host_generator(profile, cov, first, n):
    contig_len = 1_700_000
    h = synth.generate(profile, contig_len, cov, seed=7, first_region=first, num_regions=n, threads=2)
    g = synth_device.generate(profile, contig_len, cov, seed=7, first_region=first, num_regions=n, quals=True)
    torch.cuda.synchronize()
    assert g.host.n_reads == h.n_reads and g.host.n_ops == h.n_ops and g.n_bases == h.n_bases
    for name in ("read_pos", "read_base_off", "read_len", "read_cigar_off", "read_n_ops", "read_flags", "read_mapq",
                 "region_ref_start", "region_ref_end", "region_cand_start", "region_cand_end", "region_ref_off", "region_ref_len",
                 "region_read_begin", "ref"):
        got = g.t[name].cpu().numpy()[:getattr(h, name).shape[0]]
        assert np.array_equal(got, getattr(h, name)), name
    assert np.array_equal(g.t["cigar"].cpu().numpy().view(np.uint32)[:h.n_ops], h.cigar)
    assert np.array_equal(g.t["bases"].cpu().numpy()[:h.n_bases], h.bases)
    assert np.array_equal(g.t["quals"].cpu().numpy()[:h.n_bases], h.quals)
    assert g.candidate_bp == h.candidate_bp and g.read_bases == int(h.read_len.astype(np.int64).sum())


def test_generated_batch_without_qualities_runs_the_summary():
    """quals=False: PvReadBatch.quals == NULL + the profile's min_qual promise; candidates equal those of the host batch."""
    thr = synth.PROFILES["ont_r9"].thresholds
    h = synth.generate("ont_r9", 900_000, 14.0, seed=3, first_region=1, num_regions=3, threads=2)
    g = synth_device.generate("ont_r9", 900_000, 14.0, seed=3, first_region=1, num_regions=3)
    assert g.struct.quals in (0, None) and g.struct.min_qual == 5
    ws = dev.SummaryWorkspace(g.host.n_reads, g.host.n_ops, 3, g.total_positions, 1 << 14, max_region_len=int(g.region_len.max()))
    dev.summary_regions(g, thr, ws)
    k = int(ws.count.item())
    db = dev.DeviceBatch(h)
    ws2 = dev.SummaryWorkspace.for_batch(db, 1 << 14)
    dev.summary_regions(db, thr, ws2)
    assert k == int(ws2.count.item()) and k > 100 and ws.status() == 0
    assert torch.equal(ws.position[:k], ws2.position[:k]) and torch.equal(ws.windows[:k], ws2.windows[:k])
    assert torch.equal(ws.allele[:k], ws2.allele[:k])
