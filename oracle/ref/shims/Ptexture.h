/* Ptex is out of scope (Moana textures only); empty shim so headers parse. */
#pragma once
