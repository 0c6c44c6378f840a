"""Screen-space sharding of one frame across the GPUs of a node (no reference analogue).

Rank r traces (direct_stage + indirect_stage) the full-resolution row band [r*B, (r+1)*B), B = rows per rank rounded
up to a multiple of 8 (the direct stage works in 8x8 pixel tiles; the quarter-res stage lays its 8x8 tiles over ABSOLUTE tile
rows and masks the rows of a tile that belong to the neighbour, so the shared multi-bounce flag of indirect_stage.comp:283-288
stays a function of the absolute tile): 1080 rows on 8 ranks = 7 bands of 136 rows + one of 128.  Every image is allocated with world*B rows, which makes each
rank's slice of every exchange buffer the same size, so ONE exchange step of equal-sized all-gathers (in place:
rank r's slice already sits at offset r*chunk of the full buffer) gives every rank the complete pre-denoise frame.
Denoise + compose then run on the full frame on every rank; every rank ends with the complete composed image.
"""


def band_rows(height, world):
    return ((height + world - 1) // world + 7) // 8 * 8


def padded_height(height, world):
    return band_rows(height, world) * world if world > 1 else height


def band_range(rank, world, height):
    b = band_rows(height, world)
    alloc = padded_height(height, world)
    return min(rank * b, alloc), min((rank + 1) * b, alloc)


def all_gather_bands(dist, full, rank, world, inplace=True):
    """All-gather a row-banded buffer: `full` is a flat byte tensor whose length is world * chunk; rank r owns chunk r."""
    chunk = full.numel() // world
    assert chunk * world == full.numel()
    mine = full[rank * chunk:(rank + 1) * chunk]
    if not inplace:
        mine = mine.clone()
    dist.all_gather_into_tensor(full, mine)
    return full


def stripe_layout(height, world, groups=2):
    """Interleaved ownership (load balance: image regions differ a lot in tracing cost).  Rows are cut into stripes of S rows
    (multiple of 16); stripe i belongs to rank i % world; `world` consecutive stripes form one exchange group.
    Returns (S, padded allocation height = S * world * groups)."""
    if world == 1:
        return (height + 15) // 16 * 16, height
    s = ((height + world * groups - 1) // (world * groups) + 15) // 16 * 16
    return s, s * world * groups


def stripes_of(rank, world, height, groups=2):
    s, alloc = stripe_layout(height, world, groups)
    return [(g * world * s + rank * s, g * world * s + (rank + 1) * s) for g in range(groups)]
