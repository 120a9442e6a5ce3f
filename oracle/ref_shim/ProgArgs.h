/* intentionally empty: the offset generators include ProgArgs.h but use nothing from it */
#pragma once
