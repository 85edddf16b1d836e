"""List-scheduling model of a launch of one-warp tile CTAs on 148 SMs x 12 slots, to read the measured gains of the
split-tail launch (profiles/r2_split_confirm.txt, r2_split_small.txt).

Blocks are handed out in index order to whichever slot frees first (what the hardware block scheduler does).  A whole
tile takes d_sm on SM sm -- the measured first-wave means of profiles/r2_tile_times.txt, 517 ... 641 us in blocks of SM
ids -- scaled by how full the SM is when the block starts (a tile on an emptier SM runs faster: later waves measured
587 -> 566 -> 520 us); an upper piece [0, 82) takes 0.55 of a tile, a lower piece [82, 137) 0.45 (the upper levels are
the bandwidth-bound ones; the split is by measured level times, profiles/r1_level_timing.txt).

usage: python tools/model/split_tail_model.py            # the seven measured sizes, both schedules
"""
import heapq
import sys

SM_US = [517, 519, 585, 582, 568, 568, 641, 636, 635, 633, 627, 635, 624, 626, 628, 632,
         526, 521, 579, 581, 576, 567, 633, 640, 641, 641, 632, 637, 630, 631, 636, 634,
         523, 523, 583, 577, 570, 566, 584, 577, 582, 581, 562, 562, 563, 565, 575, 577,
         523, 527, 577, 576, 566, 565, 581, 577, 582, 585, 565, 565, 562, 561, 564, 568,
         520, 522, 588, 588, 568, 570, 636, 637, 629, 631, 628, 634, 630, 634, 634, 633,
         579, 583, 563, 566, 581, 583, 586, 578, 565, 569, 562, 557, 567, 570, 577, 583,
         572, 574, 581, 582, 585, 582, 565, 565, 567, 564, 569, 566, 588, 579, 571, 575,
         586, 586, 573, 576, 565, 560, 566, 568, 563, 568, 583, 584, 567, 565, 638, 631,
         632, 634, 628, 632, 619, 621, 632, 638, 559, 565, 562, 566, 569, 571, 521, 520,
         586, 576, 582, 590]
NSM, PER_SM = 148, 12
UPPER, LOWER = 0.55, 0.45
# a tile started on an SM with n of 12 slots busy (itself included) runs at this fraction of the full-SM time:
# 12 busy -> 1.0 (first wave, 585 us mean), emptier -> faster (third wave of a 2.88-round launch: 518 us = 0.885)
def occupancy_factor(n):
    return 0.80 + 0.20 * (n / 12.0)


def launch(ntiles, nsplit):
    """returns the end of the launch in us; blocks = [upper pieces][whole tiles][lower pieces]"""
    blocks = [UPPER] * nsplit + [1.0] * (ntiles - nsplit) + [LOWER] * nsplit
    busy = [0] * NSM
    free = [(0.0, sm) for sm in range(NSM) for _ in range(PER_SM)]   # (time the slot frees, sm)
    heapq.heapify(free)
    end = 0.0
    running = []            # (end time, sm) of started blocks, to keep `busy` current
    for frac in blocks:
        t, sm = heapq.heappop(free)
        while running and running[0][0] <= t:
            _, s = heapq.heappop(running)
            busy[s] -= 1
        busy[sm] += 1
        d = frac * SM_US[sm] * occupancy_factor(busy[sm])
        heapq.heappush(running, (t + d, sm))
        heapq.heappush(free, (t + d, sm))
        end = max(end, t + d)
    return end


MEASURED = {1024: (464.7, 476.9), 1792: (950.0, 796.5), 2048: (1004.0, 876.5), 2816: (1139.0, 1080.0),
            3752: (1602.0, 1391.5), 5120: (1795.0, 1706.5), 20480: (6545.0, 6429.0)}

if __name__ == "__main__":
    print("tiles  rounds | model: one CTA/tile  split tail  gain | measured: one CTA/tile  split tail  gain")
    for nt, (m0, m3) in MEASURED.items():
        a = launch(nt, 0)
        b = launch(nt, min(5 * NSM, nt))
        print(f"{nt:6d} {nt / (NSM * PER_SM):6.2f} | {a:10.0f} us {b:10.0f} us  x{a / b:.3f} | {m0:10.0f} us {m3:10.0f} us  x{m0 / m3:.3f}")
    if "--sweep" in sys.argv:
        for per in (0, 2, 3, 4, 5, 6, 8, 12):
            print(f"5120 tiles, {per:2d} split per SM: {launch(5120, per * NSM):.0f} us")
