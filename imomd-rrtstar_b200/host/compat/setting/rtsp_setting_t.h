// host/compat/setting/rtsp_setting_t.h -- ECI-Gen settings, field-compatible with the
// reference's rtsp_setting_t (include/setting/rtsp_setting_t.h:37-46 upstream).
// Put host/compat on the include path only when the reference's include tree is not.
#pragma once

typedef struct rtsp_setting
{
    bool swapping;          // try the swap variants of the cheapest insertion
    bool genetic;           // run the genetic refinement
    int ga_mutation_iter;   // seeding mutations
    int ga_population;      // offspring per generation
    int ga_generation;      // generations
    bool ga_random_seed;    // seed the solver's mt19937 from random_device
} rtsp_setting_t;

