// stands in for leica_scanstation_utils/LeicaUtils.h (scanner file paths; nothing the tests call)
#pragma once
