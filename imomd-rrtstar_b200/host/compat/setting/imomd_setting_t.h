// host/compat/setting/imomd_setting_t.h -- planner settings, field-compatible with the
// reference's imomd_setting_t (include/setting/imomd_setting_t.h:39-48 upstream).
// Put host/compat on the include path only when the reference's include tree is not.
#pragma once

#include "rtsp_setting_t.h"

typedef struct imomd_setting
{
    double goal_bias;
    int max_iter;
    int max_time;
    bool random_seed;
    bool pseudo_mode;
    int log_data;
    rtsp_setting_t rtsp_setting;
} imomd_setting_t;

