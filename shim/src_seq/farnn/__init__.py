from src_seq import chain

chain(__path__, 'farnn')
