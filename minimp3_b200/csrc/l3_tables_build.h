/*
 * l3_tables_build.h -- host-side construction of the transform kernel's constant tables from closed
 * forms: IMDCT / antialias / DCT twiddles, the short-block reorder permutation and the per-phase
 * synthesis-window taps.  Plain C++ so the CPU unit tests can build the very same tables.
 */
#ifndef MP3B_L3_TABLES_BUILD_H
#define MP3B_L3_TABLES_BUILD_H
#include <math.h>
#include <stdint.h>

#include "l3_math.cuh"
#include "mp3_tables.h"

namespace mp3b {

inline void build_consts(ImdctConsts *k)
{
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < 9; i++)
    {
        const double ang = (42.5 - 5.0*i)*pi/180.0;
        k->twid9[i] = (float)cos(ang);
        k->twid9[9 + i] = (float)sin(ang);
        /* normal window, folded: win[i] = sin(pi/36 (18.5 + i)) , win[9+i] = sin(pi/36 (i + .5)) */
        k->win[0][i] = (float)cos(pi/36.0*(i + 0.5));
        k->win[0][9 + i] = (float)sin(pi/36.0*(i + 0.5));
        /* stop block (type 3) first half: 0 x6, short sine, 1 x6 -- in the same folded layout */
        k->win[1][i] = i < 6 ? 1.f : (float)cos(pi/12.0*(i - 6 + 0.5));
        k->win[1][9 + i] = i < 6 ? 0.f : (float)sin(pi/12.0*(i - 6 + 0.5));
    }
    for (int i = 0; i < 3; i++)
    {
        const double ang = (37.5 - 15.0*i)*pi/180.0;
        k->twid3[i] = (float)cos(ang);
        k->twid3[3 + i] = (float)sin(ang);
    }
    static const double ci[8] = { -0.6, -0.535, -0.33, -0.185, -0.095, -0.041, -0.0142, -0.0037 };
    for (int i = 0; i < 8; i++)
    {
        const double n = sqrt(1.0 + ci[i]*ci[i]);
        k->aa_cs[i] = (float)(1.0/n);
        k->aa_ca[i] = (float)(-ci[i]/n);
    }
    int o = 0;
    for (int n = 32; n >= 2; n >>= 1)
        for (int i = 0; i < n/2; i++) k->dct_sec[o++] = (float)(1.0/(2.0*cos(pi*(2*i + 1)/(2.0*n))));
}

inline void build_reorder_map(uint16_t *map /* [16][576] */)
{
    for (int sr = 0; sr < 8; sr++)
        for (int mixed = 0; mixed < 2; mixed++)
        {
            uint16_t *m = map + (sr*2 + mixed)*576;
            for (int e = 0; e < 576; e++) m[e] = (uint16_t)e;
            const uint8_t *sfb;
            int start;
            if (mixed)
            {
                /* long part: the first n_long_sfb bands of the mixed table; the kernel treats
                 * n_long_bands = 2 (4 at 8 kHz) subbands as long (minimp3.h:1260) */
                const int n_long_sfb = sr >= 5 ? 8 : 6;
                sfb = MP3T_SFB_MIXED + sr*40 + n_long_sfb;
                start = (sr == 1 ? 4 : 2)*18;
            } else
            {
                sfb = MP3T_SFB_SHORT + sr*40;
                start = 0;
            }
            int src = start, dst = start;
            for (; *sfb && src < 576; sfb += 3)
            {
                const int len = *sfb;
                for (int i = 0; i < len && src + i + 2*len < 576; i++)
                {
                    m[src + i] = (uint16_t)(dst + 3*i);
                    m[src + i + len] = (uint16_t)(dst + 3*i + 1);
                    m[src + i + 2*len] = (uint16_t)(dst + 3*i + 2);
                }
                src += 3*len;
                dst += 3*len;
            }
        }
}

inline void build_fir(float *coef /* [32][16] */)
{
    for (int j = 0; j < 32; j++)
        for (int m = 0; m < 8; m++)
        {
            const float sa = j < 16 ? 1.f : (j == 16 ? 0.f : -1.f);
            coef[j*16 + 2*m] = sa*(float)MP3T_DWIN[64*m + j];
            coef[j*16 + 2*m + 1] = -(float)MP3T_DWIN[64*m + 32 + j];
        }
}

}  // namespace mp3b
#endif
